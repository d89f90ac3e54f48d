// vsm_api.cu -- C-ABI (include/vsmom_b200.h): context, workspace planning, launches.
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/vsmom_b200.h"
#include "vsm_launch.cuh"
#include "vsm_l0.cuh"
#include "vsm_stream.cuh"
#include "vsm_lin.cuh"
#include "vsm_raman_args.h"
#include "vsm_ss.cuh"

using namespace vsm;

static thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct vsm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    int max_smem_optin = 0;
    uint64_t ws_limit = 0;
    int inverse_mode = 1;   // 1: adaptive Neumann / Gauss-Jordan, 0: Gauss-Jordan only
    uint64_t launches = 0;
    int dbg_layer_ctas = 0;                 // VSM_DEBUG_LAYER_CTAS_PER_SM, read once at vsm_create
    std::string err;
    // grow-only scratch
    DevBuf consts, props, added, slabs, partial, counter, l0tmp, geom, surfpad, linprops, rrs;
    DevBuf rrsunits;                        // (n1, dn) unit list of the Raman kernels, cached across calls with the same key
    std::vector<int> rrs_key;               // nBand, n1lo, n1cnt, shifts...
    size_t rrs_nunits = 0;
    DevBuf padtab;                          // zero-padded Z / surface tables of an embedded column (see embed_size)
    int embed_min = 17;                     // smallest N that runs embedded in the N = 45 kernels (VSM_EMBED_MIN); measured: even N = 17
                                            // is 2.05x faster padded to 45 than on the run-time-N SIMT kernels
    int use_embed = 1;                      // run N between the specialised sizes as the next one, zero-padded (VSM_DISABLE_EMBED=1: off)
    int use_small = 1;                      // warp-per-point column kernel for N <= 16 in fp64 (VSM_DISABLE_SMALL=1 turns it off)
    int rrs_order = 1;                      // unit enumeration of the Raman kernels (VSM_RRS_ORDER): 1 donor-major over a host-built list of
                                            // the owner range's valid units (consecutive warps share the donor records: +7 % on the full
                                            // band), 0 dn-major
    int use_stream = 1;                     // streamed-operand kernels for N = 60 / 75 / 100 (VSM_DISABLE_STREAM=1 turns them off)
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float t_total = 0, t_layer = 0, t_add = 0;
    bool timing_valid = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> seg_layer, seg_add;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    // pinned staging of the per-call constant blob (so that the upload never blocks the host)
    void* stage = nullptr;
    size_t stage_cap = 0;
    cudaEvent_t stage_done = nullptr;
};

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);        \
            return VSM_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (device, kernel) in the PROCESS, not of a context: it is
// raised here and never lowered -- a second context asking for less (fewer layers, say) must not shrink what a first
// context was granted, or the first one's next launch fails with "invalid argument".  Issued once per growth, not per call.
#include <mutex>
static cudaError_t set_smem(vsm_ctx* ctx, const void* kern, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> granted;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(ctx->device, kern);
    auto it = granted.find(key);
    if (it != granted.end() && it->second >= (int)bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) granted[key] = (int)bytes;
    return e;
}

static int fail(vsm_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

static int ensure(vsm_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return VSM_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        char buf[128];
        snprintf(buf, sizeof buf, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return fail(ctx, VSM_ERR_NOMEM, buf);
    }
    b.cap = want;
    return VSM_OK;
}

// nullptr (and ctx->err set) when the event cannot be created
static cudaEvent_t get_event(vsm_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        const cudaError_t rc = cudaEventCreate(&e);
        if (rc != cudaSuccess) {
            ctx->err = std::string("cudaEventCreate: ") + cudaGetErrorString(rc);
            return nullptr;
        }
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}
#define GET_EVENT(var)                                  \
    cudaEvent_t var = get_event(ctx);                   \
    if (!var) return VSM_ERR_CUDA

extern "C" int vsm_version(void) { return 100; }

extern "C" int vsm_create(vsm_ctx** out, int device, void* cuda_stream) {
    if (!out) return fail(nullptr, VSM_ERR_INVALID_ARG, "vsm_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, VSM_ERR_CUDA,
                    std::string("vsm_create: no CUDA device (there is no CPU fallback): ") +
                        cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, VSM_ERR_INVALID_ARG, "vsm_create: bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, VSM_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, VSM_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, VSM_ERR_UNSUPPORTED, "vsm_create: this library is built for sm_100a (Blackwell B200) only");
    vsm_ctx* ctx = new vsm_ctx();
    ctx->device = device;
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->ws_limit = 0;
    if (const char* e = getenv("VSM_DEBUG_LAYER_CTAS_PER_SM")) ctx->dbg_layer_ctas = atoi(e);
    if (const char* e = getenv("VSM_DISABLE_STREAM")) ctx->use_stream = atoi(e) ? 0 : 1;
    if (const char* e = getenv("VSM_DISABLE_EMBED")) ctx->use_embed = atoi(e) ? 0 : 1;
    if (const char* e = getenv("VSM_EMBED_MIN")) ctx->embed_min = std::max(17, atoi(e));
    if (const char* e = getenv("VSM_DISABLE_SMALL")) ctx->use_small = atoi(e) ? 0 : 1;
    if (const char* e = getenv("VSM_RRS_ORDER")) ctx->rrs_order = atoi(e) ? 1 : 0;
    *out = ctx;
    return VSM_OK;
}

extern "C" int vsm_destroy(vsm_ctx* ctx) {
    if (!ctx) return VSM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->consts, &ctx->props, &ctx->added, &ctx->slabs, &ctx->partial, &ctx->counter, &ctx->l0tmp, &ctx->geom, &ctx->surfpad, &ctx->linprops,
                      &ctx->rrs, &ctx->rrsunits, &ctx->padtab};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    if (ctx->stage_done) cudaEventDestroy(ctx->stage_done);
    delete ctx;
    return VSM_OK;
}

extern "C" const char* vsm_last_error(vsm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int vsm_synchronize(vsm_ctx* ctx) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return VSM_OK;
}

extern "C" int vsm_set_workspace_limit(vsm_ctx* ctx, uint64_t bytes) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    ctx->ws_limit = bytes;
    return VSM_OK;
}

extern "C" int vsm_set_inverse_mode(vsm_ctx* ctx, int mode) {
    if (!ctx || (mode != 0 && mode != 1)) return VSM_ERR_INVALID_ARG;
    ctx->inverse_mode = mode;
    return VSM_OK;
}

extern "C" uint64_t vsm_kernel_launches(vsm_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------
// FMA peak microbenchmark (roofline denominator for the fp64 / fp32 SIMT path)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T seed) {
    T a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = T(0.999), c = T(0.001);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == T(123.456)) out[0] = a0;
}

template <typename T>
static int fma_peak_impl(vsm_ctx* ctx, double* tflops) {
    CK(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->l0tmp, 256);
    if (rc) return rc;
    const int iters = sizeof(T) == 8 ? 4096 : 16384;
    const int grid = ctx->num_sms * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        k_fma_peak<T><<<grid, 256, 0, ctx->stream>>>((T*)ctx->l0tmp.p, iters, T(1));
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return VSM_OK;
}

extern "C" int vsm_measure_fma_peak(vsm_ctx* ctx, int kind, double* tflops) {
    if (!ctx || !tflops) return VSM_ERR_INVALID_ARG;
    return kind == 0 ? fma_peak_impl<double>(ctx, tflops) : fma_peak_impl<float>(ctx, tflops);
}

// ---------------------------------------------------------------------------
// L0
// ---------------------------------------------------------------------------
template <typename T>
static int bmm_impl(vsm_ctx* ctx, T* C, const T* A, const T* B, int M, int K, int Nc, int batch) {
    if (!ctx || !C || !A || !B) return fail(ctx, VSM_ERR_INVALID_ARG, "batched_mul: NULL argument");
    if (M <= 0 || K <= 0 || Nc <= 0 || batch < 0) return fail(ctx, VSM_ERR_INVALID_ARG, "batched_mul: bad dims");
    if (batch == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    if (sizeof(T) == 8 && M == kL0N && K == kL0N && Nc == kL0N) {          // the hot-path shape: tensor-core path
        // (+2 KB: the fragment loads of the padded 48x48 tiles read up to 3 columns = 1.1 KB past the end of
        //  an operand; harmless inside the arena of the fused kernels, but the last stage needs the room here)
        const size_t smem = (size_t)4 * kL0LB * sizeof(double) + 2048;
        CK(set_smem(ctx, (const void*)k_bmm_f64_n45, smem));
        k_bmm_f64_n45<<<std::min(batch, ctx->num_sms * 3), 128, smem, ctx->stream>>>((double*)C, (const double*)A,
                                                                                    (const double*)B, batch);
        ctx->launches++;
        CK(cudaGetLastError());
        return VSM_OK;
    }
    const size_t sm = ((size_t)(M | 1) * K + (size_t)(K | 1) * Nc) * sizeof(T);
    const int stage = sm <= (size_t)ctx->max_smem_optin ? 1 : 0;
    const size_t smem = stage ? sm : 0;
    const int grid = std::min(batch, ctx->num_sms * 8);
    const bool vec = (Nc == 1);
    if (vec) {
        auto kern = k_bmm<T, 1, 1>;
        CK(set_smem(ctx, (const void*)kern, smem));
        kern<<<grid, 128, smem, ctx->stream>>>(C, A, B, M, K, Nc, batch, stage);
    } else if (M <= 24) {
        auto kern = k_bmm<T, 2, 2>;
        CK(set_smem(ctx, (const void*)kern, smem));
        kern<<<grid, 128, smem, ctx->stream>>>(C, A, B, M, K, Nc, batch, stage);
    } else {
        auto kern = k_bmm<T, 4, 4>;
        CK(set_smem(ctx, (const void*)kern, smem));
        kern<<<grid, 256, smem, ctx->stream>>>(C, A, B, M, K, Nc, batch, stage);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return VSM_OK;
}

template <typename T>
static int binv_impl(vsm_ctx* ctx, T* X, T* A, int N, int batch, int* info) {
    if (!ctx || !X || !A) return fail(ctx, VSM_ERR_INVALID_ARG, "batch_inv: NULL argument");
    if (N <= 0 || batch < 0) return fail(ctx, VSM_ERR_INVALID_ARG, "batch_inv: bad dims");
    if (batch == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    if (sizeof(T) == 8 && N == kL0N) {                                     // the hot-path shape: tensor-core path
        const size_t smem = ((size_t)kL0LB + 13 * kL0N) * sizeof(double) + (kL0N + 4) * sizeof(int);
        CK(set_smem(ctx, (const void*)k_binv_f64_n45, smem));
        k_binv_f64_n45<<<std::min(batch, ctx->num_sms * 3), 128, smem, ctx->stream>>>((double*)X, (const double*)A, batch, info);
        ctx->launches++;
        CK(cudaGetLastError());
        return VSM_OK;
    }
    const size_t vecb = (size_t)3 * N * sizeof(T) + (size_t)N * sizeof(int);
    const size_t full = (size_t)(N | 1) * N * sizeof(T) + vecb;
    const int stage = full <= (size_t)ctx->max_smem_optin ? 1 : 0;
    const size_t smem = stage ? full : vecb;
    auto kern = k_binv<T>;
    CK(set_smem(ctx, (const void*)kern, smem));
    const int grid = std::min(batch, ctx->num_sms * 8);
    const int threads = N <= 8 ? 64 : (N <= 24 ? 128 : 256);
    kern<<<grid, threads, smem, ctx->stream>>>(X, A, N, batch, info, stage);
    ctx->launches++;
    CK(cudaGetLastError());
    return VSM_OK;
}

template <typename T>
static int bsolve_impl(vsm_ctx* ctx, T* X, T* A, const T* B, int N, int Nc, int batch) {
    if (!ctx || !X || !A || !B) return fail(ctx, VSM_ERR_INVALID_ARG, "batch_solve: NULL argument");
    if (N <= 0 || Nc <= 0 || batch < 0) return fail(ctx, VSM_ERR_INVALID_ARG, "batch_solve: bad dims");
    if (batch == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->l0tmp, (size_t)N * N * batch * sizeof(T));
    if (rc) return rc;
    T* inv = (T*)ctx->l0tmp.p;
    rc = binv_impl<T>(ctx, inv, A, N, batch, nullptr);
    if (rc) return rc;
    return bmm_impl<T>(ctx, X, inv, B, N, N, Nc, batch);
}

extern "C" int vsm_batched_mul_f64(vsm_ctx* c, double* C, const double* A, const double* B, int M, int K, int Nc, int batch) { return bmm_impl<double>(c, C, A, B, M, K, Nc, batch); }
extern "C" int vsm_batched_mul_f32(vsm_ctx* c, float* C, const float* A, const float* B, int M, int K, int Nc, int batch) { return bmm_impl<float>(c, C, A, B, M, K, Nc, batch); }
extern "C" int vsm_batch_inv_f64(vsm_ctx* c, double* X, double* A, int N, int batch, int* info) { return binv_impl<double>(c, X, A, N, batch, info); }
extern "C" int vsm_batch_inv_f32(vsm_ctx* c, float* X, float* A, int N, int batch, int* info) { return binv_impl<float>(c, X, A, N, batch, info); }
extern "C" int vsm_batch_solve_f64(vsm_ctx* c, double* X, double* A, const double* B, int N, int Nc, int batch) { return bsolve_impl<double>(c, X, A, B, N, Nc, batch); }
extern "C" int vsm_batch_solve_f32(vsm_ctx* c, float* X, float* A, const float* B, int N, int Nc, int batch) { return bsolve_impl<float>(c, X, A, B, N, Nc, batch); }

// ---------------------------------------------------------------------------
// kernel launch helpers (tile size is a compile-time parameter chosen from N)
// ---------------------------------------------------------------------------
struct KernelCfg {
    int threads;
    int tm;
};
static KernelCfg pick_cfg(int N) {
    if (N <= 8) return {64, 1};
    if (N <= 24) return {128, 2};
    if (N <= 48) return {256, 3};
    return {256, 4};
}

template <typename T>
struct Plan {
    int ld;
    size_t msz_bytes;
    int layer_smem_mats, add_smem_mats;
    size_t layer_smem, add_smem;
    int layer_grid_per_sm, add_grid_per_sm;
};

template <typename T>
static size_t layer_vec_bytes(int n) { return ((size_t)kVecUnits * n + kMaxComp) * sizeof(T) + ((size_t)2 * n + 4) * 4; }
template <typename T>
static size_t add_vec_bytes(int n) { return ((size_t)kVecUnits * n) * sizeof(T) + ((size_t)2 * n + 4) * 4; }

// specialised (compile-time N) instantiations that exist in this build
static bool has_fast(int N) { return N == 45; }

template <typename T>
static int do_launch(vsm_ctx* ctx, bool layer, int tm, bool fast, int N, const ColArgs<T>& a, unsigned total, int grid,
                     int threads, size_t smem) {
    cudaError_t e = cudaSuccess;
#define VSM_L(TM_, NN_) (layer ? launch_layer<T, TM_, NN_>(ctx->stream, a, total, grid, threads, smem) \
                               : launch_adding<T, TM_, NN_>(ctx->stream, a, total, grid, threads, smem))
    if (fast && N == 45) e = VSM_L(0, 45);
    else switch (tm) {
        case 1: e = VSM_L(1, 0); break;
        case 2: e = VSM_L(2, 0); break;
        case 3: e = VSM_L(3, 0); break;
        default: e = VSM_L(4, 0); break;
    }
#undef VSM_L
    ctx->launches++;
    if (e != cudaSuccess) {
        ctx->err = std::string(layer ? "k_layer launch: " : "k_adding launch: ") + cudaGetErrorString(e);
        return VSM_ERR_CUDA;
    }
    return VSM_OK;
}

// the streamed kernels when a descriptor is given (fp64 only), else the shared-memory-arena kernels
template <typename T>
static int launch_col(vsm_ctx* ctx, const StreamDesc* sd, bool layer, int tm, bool fast, int N, const ColArgs<T>& a,
                      unsigned total, int grid, int threads, size_t smem) {
    if constexpr (std::is_same<T, double>::value) {
        if (sd) {
            const cudaError_t e = layer ? sd->launch_layer(ctx->stream, a, total, grid) : sd->launch_adding(ctx->stream, a, total, grid);
            ctx->launches++;
            if (e != cudaSuccess) {
                ctx->err = std::string(layer ? "k_layer_stream launch: " : "k_adding_stream launch: ") + cudaGetErrorString(e);
                return VSM_ERR_CUDA;
            }
            return VSM_OK;
        }
    }
    return do_launch<T>(ctx, layer, tm, fast, N, a, total, grid, threads, smem);
}

template <typename T>
static void get_kernel_ptrs(int tm, bool fast, int N, const void** kl, const void** ka) {
    if (fast && N == 45) { kernel_ptrs<T, 0, 45>(kl, ka); return; }
    switch (tm) {
        case 1: kernel_ptrs<T, 1, 0>(kl, ka); break;
        case 2: kernel_ptrs<T, 2, 0>(kl, ka); break;
        case 3: kernel_ptrs<T, 3, 0>(kl, ka); break;
        default: kernel_ptrs<T, 4, 0>(kl, ka); break;
    }
}

template <typename T>
static int occupancy(const void* kern, int threads, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        nb = 1;
    }
    return std::max(nb, 1);
}

// ---------------------------------------------------------------------------
// L2: whole column
// ---------------------------------------------------------------------------
// ---- embedding: a column whose N lies between the sizes that have tensor-core kernels runs as the next such size with
// DUMMY streams: zero quadrature weight, zero rows/columns in every Z table and surface moment.  A zero-weight stream is
// something the algorithm already has (the VZA / SZA streams, elemental.jl:325-332): r = 0 and t = exp(-dτ/μ) on its
// diagonal; with Z = 0 it neither receives from nor feeds any real stream, so the real N x N block of every operator is
// unchanged (the sums just carry extra exact zeros) and (Npad/N)^3 more flops still beat the SIMT kernels by 2-13x.
template <typename T>
static int embed_size(const vsm_ctx* ctx, int N, bool has_stream_export) {
    if (!ctx->use_embed || has_stream_export) return N;
    if (N >= ctx->embed_min && N < 45) return 45;            // fp64 (tensor-core kernels) and fp32 (compile-time-N SIMT kernels)
    if (std::is_same<T, double>::value) {
        if (ctx->use_stream) {
            if (N > 45 && N < 60) return 60;
            if (N > 60 && N < 75) return 75;
            if (N > 75 && N < 100) return 100;
        }
    }
    return N;
}

// dst (Np,Np,K) <- src (N,N,K), zero outside the leading N x N block
template <typename T>
__global__ void k_pad_tables(int N, int Np, size_t K, const T* __restrict__ src, T* __restrict__ dst) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = (size_t)Np * Np;
    if (e >= per * K) return;
    const size_t k = e / per;
    const int i = (int)(e % per) % Np, j = (int)(e % per) / Np;
    dst[e] = (i < N && j < N) ? src[(size_t)i + (size_t)j * N + (size_t)N * N * k] : T(0);
}

template <typename T>
static int rt_column_impl(vsm_ctx* ctx, const vsm_column_desc* d, T* R_SFI, T* T_SFI, T* hdr,
                          T* bhr_uw, T* bhr_dw, T* st_Rmp = nullptr, T* st_Tpp = nullptr, T* st_Jm = nullptr,
                          T* st_Jp = nullptr, bool padded = false) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    // struct_size versions the descriptor: the layout up to Zmp (Lambertian only) or the full one
    // (Lambertian only), up to surf_rho (BRDF moments), or the full one (+ Cox-Munk TMS correction)
    // + extra sources: thermal emission, surface SIF)
    if (!d || (d->struct_size != (int)sizeof(vsm_column_desc) && d->struct_size != (int)offsetof(vsm_column_desc, surf_rho) &&
               d->struct_size != (int)offsetof(vsm_column_desc, ss_correction) && d->struct_size != (int)offsetof(vsm_column_desc, thermal_B)))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: descriptor NULL or struct_size mismatch");
    const void* surf_rho = (d->struct_size >= (int)offsetof(vsm_column_desc, ss_correction) && d->surface_kind == VSM_SURFACE_BRDF_FOURIER)
                               ? d->surf_rho : nullptr;
    const double* ss_corr = (d->struct_size >= (int)offsetof(vsm_column_desc, thermal_B)) ? d->ss_correction : nullptr;
    const void* therm_B = (d->struct_size == (int)sizeof(vsm_column_desc)) ? d->thermal_B : nullptr;
    const void* sif0 = (d->struct_size == (int)sizeof(vsm_column_desc)) ? d->sif0 : nullptr;
    if (d->nSpec == 0) return VSM_OK;            // empty shard: nothing to do
    if (!R_SFI || !T_SFI) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: R_SFI/T_SFI NULL");
    const int N = d->N, nSt = d->nStokes, Nz = d->Nz, nAer = d->nAer, nV = d->nVZA, nM = d->n_moments;
    const int nSpec = d->nSpec;
    const int nOut = nV * nSt;
    if (N <= 0 || nSt < 1 || nSt > 4 || (N % nSt != 0 && !padded) || Nz <= 0 || nAer < 0 || nAer + 1 > kMaxComp ||
        nV <= 0 || nM <= 0 || nSpec < 0 || d->imu0 < 1 || d->imu0 * nSt > N)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: bad dimensions");
    if (d->surface_kind != VSM_SURFACE_LAMBERTIAN_SCALAR && d->surface_kind != VSM_SURFACE_BRDF_FOURIER)
        return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column: surface_kind must be LAMBERTIAN_SCALAR or BRDF_FOURIER");
    if (d->surface_kind == VSM_SURFACE_BRDF_FOURIER && !surf_rho)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: BRDF_FOURIER needs surf_rho (and the full descriptor)");
    if (!d->mu || !d->wt || !d->vza_weight || !d->ndoubl || !d->scatter || !d->iface ||
        !d->rayl_allzero || !d->vza_stream || !d->tau_rayl || !d->tau_abs || !d->F0 || !d->Zpp ||
        !d->Zmp || (nAer > 0 && (!d->aer_omega || !d->aer_ft || !d->tau_aer)))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: NULL input array");
    if (nSpec == 0) return VSM_OK;
    CK(cudaSetDevice(ctx->device));
    const int nC = nAer + 1;
    if (!padded) {
        const int Np = embed_size<T>(ctx, N, st_Rmp || st_Tpp || st_Jm || st_Jp);
        if (Np != N) {
            // zero-padded Z tables / surface moments on the device, dummy streams on the host side
            const size_t kz = (size_t)nC * nM, per = (size_t)Np * Np;
            const size_t need = (2 * kz + (surf_rho ? (size_t)nM : 0)) * per * sizeof(T);
            int rc0 = ensure(ctx, ctx->padtab, need);
            if (rc0) return rc0;
            T* zp = (T*)ctx->padtab.p; T* zm = zp + kz * per; T* rp = zm + kz * per;
            k_pad_tables<T><<<(unsigned)((kz * per + 255) / 256), 256, 0, ctx->stream>>>(N, Np, kz, (const T*)d->Zpp, zp);
            k_pad_tables<T><<<(unsigned)((kz * per + 255) / 256), 256, 0, ctx->stream>>>(N, Np, kz, (const T*)d->Zmp, zm);
            if (surf_rho) k_pad_tables<T><<<(unsigned)(((size_t)nM * per + 255) / 256), 256, 0, ctx->stream>>>(N, Np, (size_t)nM, (const T*)surf_rho, rp);
            ctx->launches += surf_rho ? 3 : 2;
            CK(cudaGetLastError());
            std::vector<double> mu2(d->mu, d->mu + N), wt2(d->wt, d->wt + N);
            for (int k = 0; k < Np - N; ++k) {
                // dummy cosines spread over (0.3, 0.9), nudged away from the real ones (only their Beer-law diagonal is used)
                double m = 0.3 + 0.6 * (k + 0.5) / (Np - N);
                for (int it = 0; it < 64; ++it) {
                    bool clash = false;
                    for (size_t q = 0; q < mu2.size(); ++q) clash = clash || fabs(mu2[q] - m) < 1e-4;
                    if (!clash) break;
                    m += 1.7e-4;
                }
                mu2.push_back(m);
                wt2.push_back(0.0);
            }
            vsm_column_desc d2;
            memset(&d2, 0, sizeof d2);
            memcpy(&d2, d, (size_t)d->struct_size);
            d2.N = Np; d2.mu = mu2.data(); d2.wt = wt2.data(); d2.Zpp = zp; d2.Zmp = zm;
            if (surf_rho) d2.surf_rho = rp;
            return rt_column_impl<T>(ctx, &d2, R_SFI, T_SFI, hdr, bhr_uw, bhr_dw, nullptr, nullptr, nullptr, nullptr, true);
        }
    }

    // ---- small constant tables -> one device blob
    std::vector<T> fT;
    std::vector<int> fI;
    auto pushT = [&](const double* p, size_t n) { size_t o = fT.size(); for (size_t i = 0; i < n; ++i) fT.push_back((T)p[i]); return o; };
    const size_t o_mu = pushT(d->mu, N), o_wt = pushT(d->wt, N);
    std::vector<double> taumod((size_t)std::max(1, nAer * Nz)), varmod((size_t)std::max(1, nAer));
    for (int c = 0; c < nAer; ++c) {
        // createAero (compEffectiveLayerProperties.jl:67-72), evaluated in FT like the reference
        const T ft = (T)d->aer_ft[c], om = (T)d->aer_omega[c];
        varmod[c] = (double)(T)((T(1) - ft) * om / (T(1) - ft * om));
        for (int iz = 0; iz < Nz; ++iz)
            taumod[c + (size_t)nAer * iz] = (double)(T)((T(1) - ft * om) * (T)d->tau_aer[c + (size_t)nAer * iz]);
    }
    const size_t o_taumod = pushT(taumod.data(), taumod.size());
    const size_t o_varmod = pushT(varmod.data(), varmod.size());
    const size_t o_vw = pushT(d->vza_weight, (size_t)nSt * nV * nM);
    const size_t o_ss = ss_corr ? pushT(ss_corr, (size_t)nOut) : 0;
    auto pushI = [&](const int32_t* p, size_t n) { size_t o = fI.size(); for (size_t i = 0; i < n; ++i) fI.push_back(p[i]); return o; };
    const size_t o_nd = pushI(d->ndoubl, Nz), o_sc = pushI(d->scatter, Nz), o_if = pushI(d->iface, Nz);
    const size_t o_rz = pushI(d->rayl_allzero, Nz);
    // heavy-first order of the scattering layers
    std::vector<int> order;
    for (int iz = 0; iz < Nz; ++iz) if (d->scatter[iz]) order.push_back(iz);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return d->ndoubl[x] > d->ndoubl[y]; });
    const int nSched = (int)order.size();
    std::vector<int> order_pad(order);
    order_pad.resize(Nz, 0);
    const size_t o_ord = pushI(order_pad.data(), Nz);
    std::vector<int> sched_idx(Nz, 0);          // compact hand-off slot of every scattering layer
    for (int k = 0; k < nSched; ++k) sched_idx[order[k]] = k;
    const size_t o_sidx = pushI(sched_idx.data(), Nz);
    const int nSl = std::max(1, nSched);
    std::vector<int> vrow(nV);
    for (int i = 0; i < nV; ++i) {
        if (d->vza_stream[i] < 1 || d->vza_stream[i] * nSt > N) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: vza_stream out of range");
        vrow[i] = (d->vza_stream[i] - 1) * nSt;
    }
    const size_t o_vrow = pushI(vrow.data(), nV);
    for (int iz = 0; iz < Nz; ++iz)
        if (d->ndoubl[iz] < 0 || d->ndoubl[iz] > 60 || d->iface[iz] < 0 || d->iface[iz] > 3)
            return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column: ndoubl/iface out of range");
    const size_t bytesT = fT.size() * sizeof(T), bytesI = fI.size() * sizeof(int);
    const size_t offI = (bytesT + 15) & ~size_t(15);
    int rc = ensure(ctx, ctx->consts, offI + bytesI);
    if (rc) return rc;
    // The previous call on this stream may still read the device blob: the copy is ordered on the
    // stream.  It goes through a pinned staging buffer owned by the context so that nothing here
    // waits for the GPU (the staging buffer is reused only after its last upload has completed).
    if (ctx->stage_cap < offI + bytesI) {
        if (ctx->stage) { CK(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->stage); ctx->stage = nullptr; ctx->stage_cap = 0; }
        CK(cudaMallocHost(&ctx->stage, (offI + bytesI) * 2));
        ctx->stage_cap = (offI + bytesI) * 2;
    }
    if (!ctx->stage_done) CK(cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming | cudaEventBlockingSync));
    else CK(cudaEventSynchronize(ctx->stage_done));
    memcpy(ctx->stage, fT.data(), bytesT);
    memcpy((char*)ctx->stage + offI, fI.data(), bytesI);
    CK(cudaMemcpyAsync(ctx->consts.p, ctx->stage, offI + bytesI, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->stage_done, ctx->stream));
    const T* cT = (const T*)ctx->consts.p;
    const int* cI = (const int*)((char*)ctx->consts.p + offI);
    if ((rc = ensure(ctx, ctx->geom, ((size_t)2 * N * N + 2 * N) * sizeof(T) + (size_t)N * N + N))) return rc;
    T* geo1 = (T*)ctx->geom.p; T* geo2 = geo1 + (size_t)N * N;
    T* gs1 = geo2 + (size_t)N * N; T* gs2 = gs1 + N;
    unsigned char* gcls = (unsigned char*)(gs2 + N);
    k_geom<T><<<(N * N + 255) / 256, 256, 0, ctx->stream>>>(N, nSt, (d->imu0 - 1) * nSt, cT + o_mu, cT + o_wt, geo1, geo2,
                                                            gs1, gs2, gcls);
    ctx->launches++;
    CK(cudaGetLastError());

    // ---- arena plan
    const KernelCfg cfg = pick_cfg(N);
    const int ld = N | 1;
    const size_t msz = (size_t)ld * N * sizeof(T);
    const size_t smax = (size_t)ctx->max_smem_optin;
    const size_t lv = layer_vec_bytes<T>(N), av = add_vec_bytes<T>(N);
    if (lv + 1024 > smax) return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column: N too large");
    int l_mats = (int)std::min<size_t>(kLayerMats, (smax - lv) / msz);
    int a_mats = (int)std::min<size_t>(kAddMats, (smax - av) / msz);
    size_t l_smem = lv + (size_t)l_mats * msz;
    size_t a_smem = av + (size_t)a_mats * msz;
    size_t l_slab = (size_t)(kLayerMats - l_mats) * msz, a_slab = (size_t)(kAddMats - a_mats) * msz;
    // streamed-operand kernels (vsm_stream.cuh): fp64, N = 60 / 75 / 100; every matrix in a zero-padded slab
    const StreamDesc* sd = (std::is_same<T, double>::value && ctx->use_stream) ? stream_desc(N) : nullptr;
    size_t mstride = (size_t)N * N;                       // elements per hand-off matrix
    if (sd) {
        l_smem = a_smem = sd->smem;
        l_slab = (size_t)kStreamLayerMats * sd->MSZ * sizeof(T);
        a_slab = (size_t)kStreamAddMats * sd->MSZ * sizeof(T);
        mstride = (size_t)sd->MSZ;
    }
    // small matrices (N <= 16, fp64): one warp per (point, moment) runs the whole column on-chip (vsm_small.cuh); the
    // per-source thermal slot and the streams export stay on the CTA-per-item kernels
    const bool small_path = std::is_same<T, double>::value && ctx->use_small && N <= 16 && !therm_B && !st_Rmp && !st_Tpp &&
                            !st_Jm && !st_Jp;
    bool fast = !sd && has_fast(N) && l_mats == kLayerMats;
    FastCfg fc{cfg.threads, cfg.threads, 0};
    if (fast) fc = fast_cfg<T, 45>();
    // the two-group adding kernel has its own arena (11 matrices)
    const size_t a2_smem = (size_t)kAdd2Mats * msz + 64 + (size_t)kAdd2VecUnits * N * sizeof(T) + ((size_t)3 * N + 8 + Nz) * 4;
    if (fast && fc.adding_two_group && a2_smem > smax) fast = false;
    if (fast && !fc.adding_two_group && a_mats != kAddMats) fast = false;

    if (fast && fc.adding_two_group) a_smem = a2_smem;
    const void* kl = nullptr; const void* ka = nullptr;
    if (sd) { kl = sd->k_layer; ka = sd->k_adding; }
    else get_kernel_ptrs<T>(cfg.tm, fast, N, &kl, &ka);
    CK(set_smem(ctx, (const void*)kl, l_smem));
    CK(set_smem(ctx, (const void*)ka, a_smem));
    // CTA sizes: the specialised kernels fix theirs at compile time
    const int l_threads = sd ? sd->threads : (fast ? fc.layer_threads : cfg.threads);
    const int a_threads = sd ? sd->threads : (fast ? fc.adding_threads : cfg.threads);
    const int l_occ = occupancy<T>(kl, l_threads, l_smem);
    const int a_occ = occupancy<T>(ka, a_threads, a_smem);
    int l_grid_max = ctx->num_sms * l_occ, a_grid_max = ctx->num_sms * a_occ;
    if (ctx->dbg_layer_ctas > 0) l_grid_max = ctx->num_sms * std::max(1, std::min(l_occ, ctx->dbg_layer_ctas));

    // ---- workspace plan: chunk over spectral points and moments
    const size_t add_item = small_path ? 16 : (size_t)nSl * ((size_t)2 * mstride + 2 * N) * sizeof(T);   // no hand-off on the small path
    const size_t props_pt = ((size_t)Nz * (2 + nC) + (Nz + 1)) * sizeof(T);
    const size_t part_pt = (size_t)3 * nM * nOut * sizeof(T);
    uint64_t limit = ctx->ws_limit ? ctx->ws_limit : (uint64_t)24 << 30;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t avail = (uint64_t)free_b + ctx->added.cap + ctx->props.cap + ctx->partial.cap + ctx->slabs.cap;
    limit = std::min<uint64_t>(limit, avail * 8 / 10);
    const uint64_t slab_bytes = (uint64_t)l_slab * l_grid_max + (uint64_t)a_slab * a_grid_max;
    if (limit < slab_bytes + add_item + props_pt + part_pt)
        return fail(ctx, VSM_ERR_NOMEM, "rt_column: workspace limit too small for one spectral point");
    const uint64_t budget = limit - slab_bytes;
    int chunk, Mb;
    {
        const uint64_t per_pt_all = (uint64_t)add_item * nM + props_pt + part_pt;
        if ((uint64_t)nSpec * per_pt_all <= budget) {
            chunk = nSpec; Mb = nM;
        } else {
            const uint64_t per_pt_1 = (uint64_t)add_item + props_pt + part_pt;
            if ((uint64_t)nSpec * per_pt_1 <= budget) {
                chunk = nSpec;
                Mb = (int)std::min<uint64_t>(nM, (budget / nSpec - props_pt - part_pt) / add_item);
                Mb = std::max(Mb, 1);
            } else {
                Mb = 1;
                chunk = (int)std::max<uint64_t>(1, budget / per_pt_1);
            }
        }
        // keep a launch no larger than needed to fill the machine many times over
        const uint64_t enough = (uint64_t)a_grid_max * 64;
        if (Mb > 1 && (uint64_t)chunk * Mb > enough) Mb = (int)std::max<uint64_t>(1, enough / chunk);
    }
    if ((rc = ensure(ctx, ctx->added, (size_t)add_item * chunk * Mb))) return rc;
    if ((rc = ensure(ctx, ctx->props, (size_t)props_pt * chunk))) return rc;
    if ((rc = ensure(ctx, ctx->partial, (size_t)part_pt * chunk + (size_t)2 * nSt * chunk * sizeof(T)))) return rc;
    if ((rc = ensure(ctx, ctx->slabs, (size_t)std::max<uint64_t>(slab_bytes, 16)))) return rc;
    if ((rc = ensure(ctx, ctx->counter, 256))) return rc;
    T* surf_pad = nullptr;
    if (sd) {
        // the padding of every slab matrix is an invariant of the streamed kernels: start from zeros
        CK(cudaMemsetAsync(ctx->slabs.p, 0, (size_t)slab_bytes, ctx->stream));
        if (surf_rho) {
            if ((rc = ensure(ctx, ctx->surfpad, (size_t)nM * sd->MSZ * sizeof(T)))) return rc;
            surf_pad = (T*)ctx->surfpad.p;
            const size_t ne = (size_t)nM * sd->MSZ;
            k_surf_pad<T><<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(N, sd->LD, sd->MSZ, nM, (const T*)surf_rho,
                                                                                cT + o_mu, cT + o_wt, surf_pad);
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }

    CK(cudaMemsetAsync((int*)ctx->counter.p + 16, 0, 4, ctx->stream));       // non-finite output counter of this call
    // timing events
    ctx->ev_used = 0;
    ctx->seg_layer.clear();
    ctx->seg_add.clear();
    GET_EVENT(ev_begin);
    GET_EVENT(ev_end);
    CK(cudaEventRecord(ev_begin, ctx->stream));

    for (int n0 = 0; n0 < nSpec; n0 += chunk) {
        const int nS = std::min(chunk, nSpec - n0);
        // --- k_props
        T* pw = (T*)ctx->props.p;
        PropsArgs<T> pa;
        pa.nS = nS; pa.Nz = Nz; pa.nAer = nAer; pa.ldspec = (size_t)nSpec;
        pa.tau_rayl = (const T*)d->tau_rayl + n0;
        pa.tau_abs = (const T*)d->tau_abs + n0;
        pa.varpi_cab = (T)d->varpi_cabannes;
        pa.aer_tau_mod = cT + o_taumod; pa.aer_varpi_mod = cT + o_varmod;
        pa.rayl_allzero = cI + o_rz;
        pa.tau = pw; pa.varpi = pw + (size_t)Nz * nS; pa.tausum = pw + (size_t)2 * Nz * nS;
        pa.W = pw + (size_t)(2 * Nz + Nz + 1) * nS;
        k_props<T><<<(nS + 127) / 128, 128, 0, ctx->stream>>>(pa);
        ctx->launches++;
        CK(cudaGetLastError());

        T* part = (T*)ctx->partial.p;
        const size_t per_m = (size_t)nS * nOut;
        T* cb_uw = part + (size_t)3 * nM * per_m;
        T* cb_dw = cb_uw + (size_t)nSt * nS;

        ColArgs<T> a;
        memset(&a, 0, sizeof a);
        a.N = N; a.nStokes = nSt; a.Nz = Nz; a.nC = nC; a.nVZA = nV; a.ld = ld;
        a.istart = (d->imu0 - 1) * nSt;
        a.nS = nS; a.n_sched_layers = nSl; a.z_spectral = 0; a.mz0 = 0; a.neumann = ctx->inverse_mode;
        a.mu0 = (T)d->mu0; a.albedo = (T)d->albedo; a.surf_rho = (const T*)surf_rho;
        a.mu = cT + o_mu; a.wt = cT + o_wt;
        a.ndoubl = cI + o_nd; a.scatter = cI + o_sc; a.iface = cI + o_if; a.layer_order = cI + o_ord;
        a.sched_idx = cI + o_sidx;
        a.Zpp = (const T*)d->Zpp; a.Zmp = (const T*)d->Zmp;
        a.p_tau = pa.tau; a.p_varpi = pa.varpi; a.p_tausum = pa.tausum; a.p_W = pa.W;
        a.F0 = (const T*)d->F0; a.F0_off = (size_t)n0;
        a.geo1 = geo1; a.geo2 = geo2; a.gs1 = gs1; a.gs2 = gs2; a.gcls = gcls;
        a.counter = (unsigned int*)ctx->counter.p;
        a.vza_row = cI + o_vrow; a.vza_w = cT + o_vw;
        a.part_R = part; a.part_T = part + (size_t)nM * per_m; a.part_H = part + (size_t)2 * nM * per_m;
        a.bhr_uw = cb_uw; a.bhr_dw = cb_dw;
        a.st_Rmp = st_Rmp; a.st_Tpp = st_Tpp; a.st_Jm = st_Jm; a.st_Jp = st_Jp;
        a.st_off = (size_t)n0; a.st_nspec = (size_t)nSpec;
        a.add_mstride = mstride; a.surf_pad = surf_pad;
        a.therm_B = (const T*)therm_B; a.sif0 = (const T*)sif0;

        for (int m0 = 0; m0 < nM; m0 += Mb) {
            const int mb = std::min(Mb, nM - m0);
            a.m0 = m0; a.Mb = mb;
            T* ad = (T*)ctx->added.p;
            const size_t slots = (size_t)mb * nSl * nS;
            a.add_r = ad; a.add_t = ad + slots * mstride; a.add_jp = ad + 2 * slots * mstride; a.add_jm = a.add_jp + slots * N;
            if (small_path) {
                if constexpr (std::is_same<T, double>::value) {
                    const unsigned total = (unsigned)((size_t)mb * nS);
                    CK(set_smem(ctx, small::small_kernel(), small::small_smem_bytes()));
                    GET_EVENT(e0);
                    GET_EVENT(e1);
                    CK(cudaEventRecord(e0, ctx->stream));
                    CK(small::launch_small(ctx->stream, a, total));
                    ctx->launches++;
                    CK(cudaEventRecord(e1, ctx->stream));
                    ctx->seg_add.push_back({e0, e1});
                }
                continue;
            }
            // --- k_layer
            if (nSched > 0) {
                const unsigned total = (unsigned)((size_t)nSched * mb * nS);
                const int grid = (int)std::min<size_t>(total, (size_t)l_grid_max);
                a.n_smem_mats = l_mats; a.slab = (T*)ctx->slabs.p; a.slab_stride = l_slab / sizeof(T);
                CK(cudaMemsetAsync(ctx->counter.p, 0, 4, ctx->stream));
                GET_EVENT(e0);
                GET_EVENT(e1);
                CK(cudaEventRecord(e0, ctx->stream));
                rc = launch_col<T>(ctx, sd, true, cfg.tm, fast, N, a, total, grid, l_threads, l_smem);
                if (rc) return rc;
                CK(cudaEventRecord(e1, ctx->stream));
                ctx->seg_layer.push_back({e0, e1});
            }
            // --- k_adding
            {
                const unsigned total = (unsigned)((size_t)mb * nS);
                const int grid = (int)std::min<size_t>(total, (size_t)a_grid_max);
                a.n_smem_mats = a_mats;
                a.slab = (T*)ctx->slabs.p + ((size_t)l_slab * l_grid_max) / sizeof(T);
                a.slab_stride = a_slab / sizeof(T);
                CK(cudaMemsetAsync(ctx->counter.p, 0, 4, ctx->stream));
                GET_EVENT(e0);
                GET_EVENT(e1);
                CK(cudaEventRecord(e0, ctx->stream));
                rc = launch_col<T>(ctx, sd, false, cfg.tm, fast, N, a, total, grid, a_threads, a_smem);
                if (rc) return rc;
                CK(cudaEventRecord(e1, ctx->stream));
                ctx->seg_add.push_back({e0, e1});
            }
        }
        // --- k_reduce into the caller's outputs (chunk window)
        k_reduce<T><<<(unsigned)((per_m + 255) / 256), 256, 0, ctx->stream>>>(
            nM, per_m, a.part_R, a.part_T, hdr ? a.part_H : (const T*)nullptr,
            R_SFI + (size_t)n0 * nOut, T_SFI + (size_t)n0 * nOut, hdr ? hdr + (size_t)n0 * nOut : (T*)nullptr,
            ss_corr ? cT + o_ss : (const T*)nullptr, pa.tausum + (size_t)Nz * nS, (T)d->mu0, nOut, (int*)ctx->counter.p + 16);
        ctx->launches++;
        CK(cudaGetLastError());
        if (bhr_uw) CK(cudaMemcpyAsync(bhr_uw + (size_t)n0 * nSt, cb_uw, (size_t)nSt * nS * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        if (bhr_dw) CK(cudaMemcpyAsync(bhr_dw + (size_t)n0 * nSt, cb_dw, (size_t)nSt * nS * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CK(cudaEventRecord(ev_end, ctx->stream));
    ctx->ev[0] = ev_begin;
    ctx->ev[1] = ev_end;
    ctx->timing_valid = true;
    return VSM_OK;
}

extern "C" int vsm_rt_column_f64(vsm_ctx* c, const vsm_column_desc* d, double* R, double* T_, double* hdr, double* uw, double* dw) { return rt_column_impl<double>(c, d, R, T_, hdr, uw, dw); }
extern "C" int vsm_rt_column_f32(vsm_ctx* c, const vsm_column_desc* d, float* R, float* T_, float* hdr, float* uw, float* dw) { return rt_column_impl<float>(c, d, R, T_, hdr, uw, dw); }
extern "C" int vsm_rt_column_streams_f64(vsm_ctx* c, const vsm_column_desc* d, double* R, double* T_, double* hdr, double* uw,
                                         double* dw, double* Rmp, double* Tpp, double* Jm, double* Jp) {
    return rt_column_impl<double>(c, d, R, T_, hdr, uw, dw, Rmp, Tpp, Jm, Jp);
}
extern "C" int vsm_rt_column_streams_f32(vsm_ctx* c, const vsm_column_desc* d, float* R, float* T_, float* hdr, float* uw,
                                         float* dw, float* Rmp, float* Tpp, float* Jm, float* Jp) {
    return rt_column_impl<float>(c, d, R, T_, hdr, uw, dw, Rmp, Tpp, Jm, Jp);
}

// ---------------------------------------------------------------------------
// L2 LinMode: forward + Jacobians (rt_run_lin.jl:102-326)
// ---------------------------------------------------------------------------
static int rt_column_lin_impl(vsm_ctx* ctx, const vsm_column_desc* d, const vsm_lin_desc* ld_, double* R_SFI, double* T_SFI,
                              double* dR, double* dT, bool padded = false) {
    using T = double;
    if (!ctx) return VSM_ERR_INVALID_ARG;
    if (!d || d->struct_size < (int)offsetof(vsm_column_desc, ss_correction))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: descriptor NULL or struct_size mismatch");
    if (!ld_ || ld_->struct_size != (int)sizeof(vsm_lin_desc) || ld_->n_params < 1 || ld_->n_params > kMaxParams)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: lin descriptor NULL / struct_size mismatch / n_params not in 1..16");
    const void* surf_rho = (d->surface_kind == VSM_SURFACE_BRDF_FOURIER) ? d->surf_rho : nullptr;
    if (d->nSpec == 0) return VSM_OK;
    if (!R_SFI || !T_SFI || !dR || !dT) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: output NULL");
    const int N = d->N, nSt = d->nStokes, Nz = d->Nz, nAer = d->nAer, nV = d->nVZA, nM = d->n_moments, nSpec = d->nSpec;
    const int P = ld_->n_params, nOut = nV * nSt, nC = nAer + 1;
    if (N <= 0 || nSt < 1 || nSt > 4 || (N % nSt != 0 && !padded) || Nz <= 0 || nAer < 0 || nC > kMaxComp || nV <= 0 || nM <= 0 ||
        d->imu0 < 1 || d->imu0 * nSt > N)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: bad dimensions");
    if (d->surface_kind == VSM_SURFACE_BRDF_FOURIER && !surf_rho) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: BRDF_FOURIER needs surf_rho");
    if (!d->mu || !d->wt || !d->vza_weight || !d->ndoubl || !d->scatter || !d->iface || !d->vza_stream || !d->tau_rayl ||
        !d->tau_abs || !d->F0 || !d->Zpp || !d->Zmp || (nAer > 0 && (!d->aer_omega || !d->aer_ft || !d->tau_aer)))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: NULL input array");
    for (int iz = 0; iz < Nz; ++iz) {
        if (!d->scatter[iz] || (iz > 0 && d->iface[iz] != VSM_IFACE_11))
            return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_lin: every layer must scatter (ScatteringInterface_11 throughout)");
        if (d->ndoubl[iz] < 0 || d->ndoubl[iz] > 60) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: ndoubl out of range");
    }
    CK(cudaSetDevice(ctx->device));
    if (!padded && ctx->use_embed && ctx->use_stream) {
        // embedding (see embed_size): the LinMode tensor-core kernels exist for N = 60 / 75 / 100; a smaller N runs as the next
        // of them with dummy zero-weight streams (the derivative Z tables are mixed on the device from the same padded base
        // tables), measured 2.0x over the run-time-N SIMT LinMode kernels at N = 45; taken from N = 40 on ((60/40)^3 = 3.4x the flops)
        const int Np = (N >= 40 && N < 60) ? 60 : (N > 60 && N < 75) ? 75 : (N > 75 && N < 100) ? 100 : N;
        if (Np != N) {
            const size_t kz = (size_t)nC * nM, per = (size_t)Np * Np;
            const size_t need = (2 * kz + (surf_rho ? (size_t)nM : 0)) * per * sizeof(T);
            int rc0 = ensure(ctx, ctx->padtab, need);
            if (rc0) return rc0;
            T* zp = (T*)ctx->padtab.p; T* zm = zp + kz * per; T* rp = zm + kz * per;
            k_pad_tables<T><<<(unsigned)((kz * per + 255) / 256), 256, 0, ctx->stream>>>(N, Np, kz, (const T*)d->Zpp, zp);
            k_pad_tables<T><<<(unsigned)((kz * per + 255) / 256), 256, 0, ctx->stream>>>(N, Np, kz, (const T*)d->Zmp, zm);
            if (surf_rho) k_pad_tables<T><<<(unsigned)(((size_t)nM * per + 255) / 256), 256, 0, ctx->stream>>>(N, Np, (size_t)nM, (const T*)surf_rho, rp);
            ctx->launches += surf_rho ? 3 : 2;
            CK(cudaGetLastError());
            std::vector<double> mu2(d->mu, d->mu + N), wt2(d->wt, d->wt + N);
            for (int k = 0; k < Np - N; ++k) {
                double mk = 0.3 + 0.6 * (k + 0.5) / (Np - N);
                for (int it = 0; it < 64; ++it) {
                    bool clash = false;
                    for (size_t q = 0; q < mu2.size(); ++q) clash = clash || fabs(mu2[q] - mk) < 1e-4;
                    if (!clash) break;
                    mk += 1.7e-4;
                }
                mu2.push_back(mk);
                wt2.push_back(0.0);
            }
            vsm_column_desc d2;
            memset(&d2, 0, sizeof d2);
            memcpy(&d2, d, (size_t)std::min<int>(d->struct_size, (int)sizeof d2));
            d2.N = Np; d2.mu = mu2.data(); d2.wt = wt2.data(); d2.Zpp = zp; d2.Zmp = zm;
            if (surf_rho) d2.surf_rho = rp;
            return rt_column_lin_impl(ctx, &d2, ld_, R_SFI, T_SFI, dR, dT, true);
        }
    }
    // ---- constant tables (as rt_column_impl) + the parameter tables
    std::vector<T> fT;
    std::vector<int> fI;
    auto pushT = [&](const double* p, size_t n) { size_t o = fT.size(); for (size_t i = 0; i < n; ++i) fT.push_back((T)p[i]); return o; };
    auto pushI = [&](const int32_t* p, size_t n) { size_t o = fI.size(); for (size_t i = 0; i < n; ++i) fI.push_back(p[i]); return o; };
    const size_t o_mu = pushT(d->mu, N), o_wt = pushT(d->wt, N);
    std::vector<double> taumod((size_t)std::max(1, nAer * Nz)), varmod((size_t)std::max(1, nAer)), dtaumod((size_t)std::max(1, nAer * Nz * P), 0.0);
    for (int c = 0; c < nAer; ++c) {
        const T ft = (T)d->aer_ft[c], om = (T)d->aer_omega[c];
        varmod[c] = (T(1) - ft) * om / (T(1) - ft * om);
        for (int iz = 0; iz < Nz; ++iz) {
            taumod[c + (size_t)nAer * iz] = (T(1) - ft * om) * (T)d->tau_aer[c + (size_t)nAer * iz];
            if (ld_->dtau_aer)
                for (int p = 0; p < P; ++p)      // d tau_mod: [(p*Nz + iz)*nAer + c] <- host (nAer,Nz,P)
                    dtaumod[((size_t)p * Nz + iz) * nAer + c] = (T(1) - ft * om) * ld_->dtau_aer[c + (size_t)nAer * (iz + (size_t)Nz * p)];
        }
    }
    const size_t o_taumod = pushT(taumod.data(), taumod.size()), o_varmod = pushT(varmod.data(), varmod.size());
    const size_t o_dtaumod = pushT(dtaumod.data(), dtaumod.size());
    std::vector<double> dalb(P, 0.0);
    if (ld_->dalbedo) for (int p = 0; p < P; ++p) dalb[p] = ld_->dalbedo[p];
    const size_t o_dalb = pushT(dalb.data(), P);
    const size_t o_vw = pushT(d->vza_weight, (size_t)nSt * nV * nM);
    const size_t o_nd = pushI(d->ndoubl, Nz), o_sc = pushI(d->scatter, Nz), o_if = pushI(d->iface, Nz);
    std::vector<int> rz(Nz, 0);
    if (d->rayl_allzero) for (int iz = 0; iz < Nz; ++iz) rz[iz] = d->rayl_allzero[iz];
    const size_t o_rz = pushI(rz.data(), Nz);
    std::vector<int> order(Nz);
    for (int iz = 0; iz < Nz; ++iz) order[iz] = iz;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return d->ndoubl[x] > d->ndoubl[y]; });
    std::vector<int> sidx(Nz);
    for (int k = 0; k < Nz; ++k) sidx[order[k]] = k;
    const size_t o_ord = pushI(order.data(), Nz), o_sidx = pushI(sidx.data(), Nz);
    std::vector<int> vrow(nV);
    for (int i = 0; i < nV; ++i) {
        if (d->vza_stream[i] < 1 || d->vza_stream[i] * nSt > N) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_lin: vza_stream out of range");
        vrow[i] = (d->vza_stream[i] - 1) * nSt;
    }
    const size_t o_vrow = pushI(vrow.data(), nV);
    const size_t bytesT = fT.size() * sizeof(T), bytesI = fI.size() * sizeof(int);
    const size_t offI = (bytesT + 15) & ~size_t(15);
    int rc = ensure(ctx, ctx->consts, offI + bytesI);
    if (rc) return rc;
    // (synchronous upload from pageable memory: this entry is not on the latency-critical path)
    std::vector<char> blob(offI + bytesI);
    memcpy(blob.data(), fT.data(), bytesT);
    memcpy(blob.data() + offI, fI.data(), bytesI);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(ctx->consts.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const T* cT = (const T*)ctx->consts.p;
    const int* cI = (const int*)((char*)ctx->consts.p + offI);

    // ---- kernels and layout
    const LinDesc* ldsc = ctx->use_stream ? lin_desc_stream(N) : nullptr;
    if (!ldsc) ldsc = lin_desc_generic(N <= 24 ? 2 : 4);
    const int mld = ldsc->ld ? ldsc->ld : (N | 1), mcols = ldsc->ld ? ldsc->ncols : N;
    const size_t msz = (size_t)mld * mcols;
    const int vs = ldsc->ld ? ldsc->ncols : N;
    const size_t smem = ldsc->smem ? ldsc->smem : ((size_t)13 * N + kMaxComp) * sizeof(T) + ((size_t)2 * N + 8) * 4;
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_lin: N too large");
    CK(set_smem(ctx, ldsc->k_layer, smem));
    CK(set_smem(ctx, ldsc->k_adding, smem));
    const int l_grid_max = ctx->num_sms * occupancy<T>(ldsc->k_layer, ldsc->threads, smem);
    const int a_grid_max = ctx->num_sms * occupancy<T>(ldsc->k_adding, ldsc->threads, smem);
    const size_t dsz = (size_t)(1 + P) * msz, dv = (size_t)(1 + P) * vs;
    const size_t l_slab = ((size_t)kLinLayerDuals * dsz + kLinSingles * msz + kLinLayerVecs * dv + (size_t)2 * (1 + P) * nSt * N + 16) * sizeof(T);
    const size_t a_slab = ((size_t)kLinAddDuals * dsz + kLinSingles * msz + kLinAddVecs * dv + 16) * sizeof(T);
    const uint64_t slab_bytes = (uint64_t)l_slab * l_grid_max + (uint64_t)a_slab * a_grid_max;

    // ---- workspace plan: one moment at a time, chunks of spectral points
    const size_t add_item = (size_t)Nz * (1 + P) * ((size_t)2 * msz + 2 * N) * sizeof(T);
    const size_t props_pt = ((size_t)Nz * (2 + nC) + (Nz + 1)) * sizeof(T);
    const size_t lprops_pt = (size_t)P * ((size_t)Nz * (2 + nC) + (Nz + 1)) * sizeof(T);
    const size_t part_pt = (size_t)2 * nM * nOut * (1 + P) * sizeof(T) + (size_t)nM * nOut * sizeof(T);
    uint64_t limit = ctx->ws_limit ? ctx->ws_limit : (uint64_t)24 << 30;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t avail = (uint64_t)free_b + ctx->added.cap + ctx->props.cap + ctx->partial.cap + ctx->slabs.cap + ctx->linprops.cap;
    limit = std::min<uint64_t>(limit, avail * 8 / 10);
    const uint64_t per_pt = (uint64_t)add_item + props_pt + lprops_pt + part_pt;
    if (limit < slab_bytes + per_pt) return fail(ctx, VSM_ERR_NOMEM, "rt_column_lin: workspace limit too small for one spectral point");
    const int chunk = (int)std::min<uint64_t>((uint64_t)nSpec, std::max<uint64_t>(1, (limit - slab_bytes) / per_pt));
    if ((rc = ensure(ctx, ctx->added, (size_t)add_item * chunk))) return rc;
    if ((rc = ensure(ctx, ctx->props, (size_t)props_pt * chunk))) return rc;
    if ((rc = ensure(ctx, ctx->linprops, (size_t)lprops_pt * chunk))) return rc;
    if ((rc = ensure(ctx, ctx->partial, (size_t)part_pt * chunk))) return rc;
    if ((rc = ensure(ctx, ctx->slabs, (size_t)std::max<uint64_t>(slab_bytes, 16)))) return rc;
    if ((rc = ensure(ctx, ctx->counter, 256))) return rc;
    CK(cudaMemsetAsync(ctx->slabs.p, 0, (size_t)slab_bytes, ctx->stream));      // zero padding is an invariant of the streamed products

    ctx->ev_used = 0; ctx->seg_layer.clear(); ctx->seg_add.clear();
    GET_EVENT(ev_begin);
    GET_EVENT(ev_end);
    CK(cudaEventRecord(ev_begin, ctx->stream));
    for (int n0 = 0; n0 < nSpec; n0 += chunk) {
        const int nS = std::min(chunk, nSpec - n0);
        T* pw = (T*)ctx->props.p;
        PropsArgs<T> pa;
        pa.nS = nS; pa.Nz = Nz; pa.nAer = nAer; pa.ldspec = (size_t)nSpec;
        pa.tau_rayl = (const T*)d->tau_rayl + n0; pa.tau_abs = (const T*)d->tau_abs + n0;
        pa.varpi_cab = (T)d->varpi_cabannes;
        pa.aer_tau_mod = cT + o_taumod; pa.aer_varpi_mod = cT + o_varmod; pa.rayl_allzero = cI + o_rz;
        pa.tau = pw; pa.varpi = pw + (size_t)Nz * nS; pa.tausum = pw + (size_t)2 * Nz * nS; pa.W = pw + (size_t)(3 * Nz + 1) * nS;
        k_props<T><<<(nS + 127) / 128, 128, 0, ctx->stream>>>(pa);
        ctx->launches++;
        CK(cudaGetLastError());
        T* lw = (T*)ctx->linprops.p;
        PropsLinArgs<T> pl;
        pl.nS = nS; pl.Nz = Nz; pl.nAer = nAer; pl.P = P; pl.ldspec = (size_t)nSpec;
        pl.tau_rayl = pa.tau_rayl; pl.tau_abs = pa.tau_abs;
        pl.dtau_rayl = ld_->dtau_rayl ? (const T*)ld_->dtau_rayl + n0 : nullptr;
        pl.dtau_abs = ld_->dtau_abs ? (const T*)ld_->dtau_abs + n0 : nullptr;
        pl.varpi_cab = pa.varpi_cab; pl.aer_tau_mod = pa.aer_tau_mod; pl.aer_varpi_mod = pa.aer_varpi_mod;
        pl.daer_tau_mod = (nAer > 0 && ld_->dtau_aer) ? cT + o_dtaumod : nullptr;
        pl.d_tau = lw; pl.d_varpi = lw + (size_t)P * Nz * nS; pl.d_tausum = lw + (size_t)2 * P * Nz * nS;
        pl.d_W = lw + ((size_t)2 * P * Nz + (size_t)P * (Nz + 1)) * nS;
        k_props_lin<T><<<(nS + 127) / 128, 128, 0, ctx->stream>>>(pl);
        ctx->launches++;
        CK(cudaGetLastError());

        T* part = (T*)ctx->partial.p;
        const size_t per_m = (size_t)nS * nOut;
        ColArgs<T> a;
        memset(&a, 0, sizeof a);
        a.N = N; a.nStokes = nSt; a.Nz = Nz; a.nC = nC; a.nVZA = nV; a.ld = mld;
        a.istart = (d->imu0 - 1) * nSt; a.nS = nS; a.n_sched_layers = Nz; a.neumann = ctx->inverse_mode;
        a.mu0 = (T)d->mu0; a.albedo = (T)d->albedo; a.surf_rho = (const T*)surf_rho;
        a.mu = cT + o_mu; a.wt = cT + o_wt;
        a.ndoubl = cI + o_nd; a.scatter = cI + o_sc; a.iface = cI + o_if; a.layer_order = cI + o_ord; a.sched_idx = cI + o_sidx;
        a.Zpp = (const T*)d->Zpp; a.Zmp = (const T*)d->Zmp;
        a.p_tau = pa.tau; a.p_varpi = pa.varpi; a.p_tausum = pa.tausum; a.p_W = pa.W;
        a.F0 = (const T*)d->F0; a.F0_off = (size_t)n0;
        a.counter = (unsigned int*)ctx->counter.p;
        a.vza_row = cI + o_vrow; a.vza_w = cT + o_vw;
        a.part_R = part; a.part_T = part + (size_t)nM * per_m; a.part_H = part + (size_t)2 * nM * per_m;
        a.add_mstride = msz;
        LinArgs<T> l;
        memset(&l, 0, sizeof l);
        l.P = P; l.msz = msz; l.vstride = vs;
        l.d_tau = pl.d_tau; l.d_varpi = pl.d_varpi; l.d_tausum = pl.d_tausum; l.d_W = pl.d_W; l.dalbedo = cT + o_dalb;
        l.part_dR = part + (size_t)3 * nM * per_m; l.part_dT = l.part_dR + (size_t)nM * per_m * P;
        T* ad = (T*)ctx->added.p;
        const size_t slots = (size_t)Nz * nS * (1 + P);
        l.add_r = ad; l.add_t = ad + slots * msz; l.add_jp = ad + 2 * slots * msz; l.add_jm = l.add_jp + slots * N;
        for (int m0 = 0; m0 < nM; ++m0) {
            a.m0 = m0; a.Mb = 1;
            {
                const unsigned total = (unsigned)((size_t)Nz * nS);
                const int grid = (int)std::min<size_t>(total, (size_t)l_grid_max);
                a.slab = (T*)ctx->slabs.p; a.slab_stride = l_slab / sizeof(T);
                CK(cudaMemsetAsync(ctx->counter.p, 0, 4, ctx->stream));
                GET_EVENT(e0);
                GET_EVENT(e1);
                CK(cudaEventRecord(e0, ctx->stream));
                const cudaError_t e = ldsc->launch_layer(ctx->stream, a, l, total, grid, smem);
                ctx->launches++;
                if (e != cudaSuccess) return fail(ctx, VSM_ERR_CUDA, std::string("k_layer_lin launch: ") + cudaGetErrorString(e));
                CK(cudaEventRecord(e1, ctx->stream));
                ctx->seg_layer.push_back({e0, e1});
            }
            {
                const unsigned total = (unsigned)nS;
                const int grid = (int)std::min<size_t>(total, (size_t)a_grid_max);
                a.slab = (T*)ctx->slabs.p + ((size_t)l_slab * l_grid_max) / sizeof(T); a.slab_stride = a_slab / sizeof(T);
                CK(cudaMemsetAsync(ctx->counter.p, 0, 4, ctx->stream));
                GET_EVENT(e0);
                GET_EVENT(e1);
                CK(cudaEventRecord(e0, ctx->stream));
                const cudaError_t e = ldsc->launch_adding(ctx->stream, a, l, total, grid, smem);
                ctx->launches++;
                if (e != cudaSuccess) return fail(ctx, VSM_ERR_CUDA, std::string("k_adding_lin launch: ") + cudaGetErrorString(e));
                CK(cudaEventRecord(e1, ctx->stream));
                ctx->seg_add.push_back({e0, e1});
            }
        }
        k_reduce<T><<<(unsigned)((per_m + 255) / 256), 256, 0, ctx->stream>>>(
            nM, per_m, a.part_R, a.part_T, (const T*)nullptr, R_SFI + (size_t)n0 * nOut, T_SFI + (size_t)n0 * nOut, (T*)nullptr,
            (const T*)nullptr, pa.tausum + (size_t)Nz * nS, (T)d->mu0, nOut);
        const size_t per_ml = per_m * P;
        k_reduce_lin<T><<<(unsigned)((per_ml + 255) / 256), 256, 0, ctx->stream>>>(nM, nS, P, nOut, (size_t)nSpec, (size_t)n0,
                                                                               l.part_dR, l.part_dT, dR, dT);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ev_end, ctx->stream));
    ctx->ev[0] = ev_begin; ctx->ev[1] = ev_end; ctx->timing_valid = true;
    return VSM_OK;
}

extern "C" int vsm_rt_column_lin_f64(vsm_ctx* c, const vsm_column_desc* d, const vsm_lin_desc* l, double* R, double* T_, double* dR,
                                     double* dT) {
    return rt_column_lin_impl(c, d, l, R, T_, dR, dT);
}

// ----------------------------------------------------------------------------------------------------------------
// L2, rotational Raman (RRS): rt_run(::RRS, model, iBand) rt_run.jl:238-539 with rt_kernel!(::RRS) rt_kernel.jl:352-391
// ----------------------------------------------------------------------------------------------------------------
static int rt_column_rrs_impl(vsm_ctx* ctx, const vsm_column_desc* d, const vsm_rrs_desc* rd, double* R_SFI, double* T_SFI,
                              double* ieR_SFI, double* ieT_SFI) {
    using T = double;
    if (!ctx) return VSM_ERR_INVALID_ARG;
    // the column descriptor with or without the trailing extra-source fields (which this driver does not take)
    if (!d || (d->struct_size != (int)sizeof(vsm_column_desc) && d->struct_size != (int)offsetof(vsm_column_desc, thermal_B)))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: column descriptor NULL or struct_size mismatch");
    const bool full_desc = d->struct_size == (int)sizeof(vsm_column_desc);
    if (!rd || rd->struct_size != (int)sizeof(vsm_rrs_desc))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: rrs descriptor NULL or struct_size mismatch");
    if (!R_SFI || !T_SFI || !ieR_SFI || !ieT_SFI) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: output NULL");
    const int N = d->N, nSt = d->nStokes, Nz = d->Nz, nAer = d->nAer, nV = d->nVZA, nM = d->n_moments;
    const int nBand = d->nSpec, nR = rd->n_raman;
    const int nOut = nV * nSt;
    if (N <= 0 || nSt < 1 || nSt > 4 || N % nSt != 0 || Nz <= 0 || nAer < 0 || nAer + 1 > kMaxComp || nV <= 0 || nM <= 0 ||
        nBand <= 0 || d->imu0 < 1 || d->imu0 * nSt > N || nR <= 0)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: bad dimensions");
    if (!d->mu || !d->wt || !d->vza_weight || !d->ndoubl || !d->scatter || !d->iface || !d->rayl_allzero || !d->vza_stream ||
        !d->tau_rayl || !d->tau_abs || !d->F0 || !d->Zpp || !d->Zmp || (nAer > 0 && (!d->aer_omega || !d->aer_ft || !d->tau_aer)) ||
        !rd->i_shift || !rd->varpi_raman || !rd->Zpp_raman || !rd->Zmp_raman)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: NULL input array");
    if ((full_desc && (d->thermal_B || d->sif0)) || d->ss_correction)
        return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_rrs: extra sources / TMS correction are not part of the RRS driver");
    if (d->surface_kind != VSM_SURFACE_LAMBERTIAN_SCALAR && d->surface_kind != VSM_SURFACE_BRDF_FOURIER)
        return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_rrs: surface_kind must be LAMBERTIAN_SCALAR or BRDF_FOURIER");
    if (d->surface_kind == VSM_SURFACE_BRDF_FOURIER && !d->surf_rho)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: BRDF_FOURIER needs surf_rho");
    int ndmax = 0;
    for (int iz = 0; iz < Nz; ++iz) {
        // rt_kernel!(::RRS) hard-codes scatter = true (rt_kernel.jl:366); only the _11 interaction exists for RRS
        if (!d->scatter[iz] || d->iface[iz] != VSM_IFACE_11)
            return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_rrs: every layer must scatter (ScatteringInterface_11)");
        if (d->ndoubl[iz] < 0 || d->ndoubl[iz] > 60) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: ndoubl out of range");
        ndmax = std::max(ndmax, d->ndoubl[iz]);
    }
    const int n1lo = rd->n1_begin, n1cnt = rd->n1_count;
    if (n1lo < 0 || n1cnt < 0 || n1lo + n1cnt > nBand) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: owner range outside the band");
    if (n1cnt == 0) return VSM_OK;
    const rrs::Desc* kd = rrs::desc(N);
    if (!kd) return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_rrs: N > 48 is not covered by the Raman kernels");
    const int NP = kd->NP;
    CK(cudaSetDevice(ctx->device));
    const int nC = nAer + 1;
    int smin = 0, smax = 0;
    for (int k = 0; k < nR; ++k) { smin = std::min(smin, rd->i_shift[k]); smax = std::max(smax, rd->i_shift[k]); }
    // elastic halo: the owner range plus every donor wavelength it can reach
    const int e0 = std::max(0, n1lo + smin), e1 = std::min(nBand, n1lo + n1cnt + smax);
    const int nE = e1 - e0;

    // ---- constant tables
    std::vector<T> fT;
    std::vector<int> fI;
    auto pushT = [&](const double* p, size_t n) { size_t o = fT.size(); for (size_t i = 0; i < n; ++i) fT.push_back(p[i]); return o; };
    auto pushI = [&](const int32_t* p, size_t n) { size_t o = fI.size(); for (size_t i = 0; i < n; ++i) fI.push_back(p[i]); return o; };
    std::vector<double> mu_pad(NP, 1.0), wt_pad(NP, 0.0), sgn(NP, 1.0);
    for (int i = 0; i < N; ++i) { mu_pad[i] = d->mu[i]; wt_pad[i] = d->wt[i]; sgn[i] = ((i % nSt) >= 2) ? -1.0 : 1.0; }
    const size_t o_mu = pushT(mu_pad.data(), NP), o_wt = pushT(wt_pad.data(), NP), o_sg = pushT(sgn.data(), NP);
    std::vector<double> taumod((size_t)std::max(1, nAer * Nz)), varmod((size_t)std::max(1, nAer));
    for (int c = 0; c < nAer; ++c) {
        const T ft = d->aer_ft[c], om = d->aer_omega[c];
        varmod[c] = (T(1) - ft) * om / (T(1) - ft * om);
        for (int iz = 0; iz < Nz; ++iz) taumod[c + (size_t)nAer * iz] = (T(1) - ft * om) * d->tau_aer[c + (size_t)nAer * iz];
    }
    const size_t o_taumod = pushT(taumod.data(), taumod.size()), o_varmod = pushT(varmod.data(), varmod.size());
    const size_t o_vw = pushT(d->vza_weight, (size_t)nSt * nV * nM);
    const size_t o_wR = pushT(rd->varpi_raman, nR);
    size_t o_rho = 0;
    const bool lambertian = d->surface_kind == VSM_SURFACE_LAMBERTIAN_SCALAR;
    if (lambertian) {
        // lambertian_surface.jl:59-70: R_surf = 2 albedo on the (I,I) element of every stream pair, m = 0 only
        std::vector<double> rho((size_t)N * N, 0.0);
        for (int i = 0; i < N; i += nSt) for (int j = 0; j < N; j += nSt) rho[i + (size_t)j * N] = 2.0 * d->albedo;
        o_rho = pushT(rho.data(), rho.size());
    }
    const size_t o_nd = pushI(d->ndoubl, Nz), o_rz = pushI(d->rayl_allzero, Nz), o_sh = pushI(rd->i_shift, nR);
    std::vector<int> vrow(nV);
    for (int i = 0; i < nV; ++i) {
        if (d->vza_stream[i] < 1 || d->vza_stream[i] * nSt > N) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_rrs: vza_stream out of range");
        vrow[i] = (d->vza_stream[i] - 1) * nSt;
    }
    const size_t o_vrow = pushI(vrow.data(), nV);
    const size_t bytesT = fT.size() * sizeof(T), bytesI = fI.size() * sizeof(int);
    const size_t offI = (bytesT + 15) & ~size_t(15);
    int rc = ensure(ctx, ctx->consts, offI + bytesI);
    if (rc) return rc;
    if (ctx->stage_cap < offI + bytesI) {
        if (ctx->stage) { CK(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->stage); ctx->stage = nullptr; ctx->stage_cap = 0; }
        CK(cudaMallocHost(&ctx->stage, (offI + bytesI) * 2));
        ctx->stage_cap = (offI + bytesI) * 2;
    }
    if (!ctx->stage_done) CK(cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming | cudaEventBlockingSync));
    else CK(cudaEventSynchronize(ctx->stage_done));
    memcpy(ctx->stage, fT.data(), bytesT);
    memcpy((char*)ctx->stage + offI, fI.data(), bytesI);
    CK(cudaMemcpyAsync(ctx->consts.p, ctx->stage, offI + bytesI, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->stage_done, ctx->stream));
    const T* cT = (const T*)ctx->consts.p;
    const int* cI = (const int*)((char*)ctx->consts.p + offI);

    // ---- workspace (doubles)
    const size_t U = (size_t)nR * n1cnt;
    auto even = [](size_t x) { return (x + 1) & ~size_t(1); };          // keep every array 16-byte aligned (vector loads)
    const size_t w_props = even(((size_t)Nz * (2 + nC) + Nz + 1) * nE), w_fsc = even((size_t)Nz * nE);
    const size_t w_cE = (size_t)nE * kd->ce, w_cI = U * kd->ci;
    const size_t w_rec0 = (size_t)std::max(1, ndmax) * nE * kd->rec0, w_rec1 = (size_t)std::max(1, ndmax) * nE * kd->rec1;
    const size_t w_i0 = (size_t)nE * kd->irec0, w_i1 = (size_t)nE * kd->irec1;
    const size_t w_uo = U * 2 * nOut;
    const size_t w_total = w_props + w_fsc + 2 * w_cE + 2 * w_cI + w_rec0 + w_rec1 + w_i0 + w_i1 + w_uo + 64;
    {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if (w_total * sizeof(T) > free_b + ctx->rrs.cap) {
            char buf[200];
            snprintf(buf, sizeof buf, "rt_column_rrs: needs %.1f GB of workspace (nRaman x n1_count = %zu units); shrink the owner range",
                     w_total * 8.0 / 1e9, U);
            return fail(ctx, VSM_ERR_NOMEM, buf);
        }
    }
    if ((rc = ensure(ctx, ctx->rrs, w_total * sizeof(T)))) return rc;
    if ((rc = ensure(ctx, ctx->counter, 256))) return rc;
    T* w = (T*)ctx->rrs.p;
    T* pw = w; w += w_props;
    T* fsc = w; w += w_fsc;
    T* cE[2] = {w, w + w_cE}; w += 2 * w_cE;
    T* cIe[2] = {w, w + w_cI}; w += 2 * w_cI;
    T* rec0 = w; w += w_rec0;
    T* rec1 = w; w += w_rec1;
    T* i0 = w; w += w_i0;
    T* i1 = w; w += w_i1;
    T* uo = w;

    CK(cudaMemsetAsync((int*)ctx->counter.p + 16, 0, 4, ctx->stream));
    const size_t out_bytes = (size_t)nOut * n1cnt * sizeof(T);
    CK(cudaMemsetAsync(R_SFI, 0, out_bytes, ctx->stream));
    CK(cudaMemsetAsync(T_SFI, 0, out_bytes, ctx->stream));
    CK(cudaMemsetAsync(ieR_SFI, 0, out_bytes, ctx->stream));
    CK(cudaMemsetAsync(ieT_SFI, 0, out_bytes, ctx->stream));
    ctx->ev_used = 0;
    ctx->seg_layer.clear();
    ctx->seg_add.clear();
    GET_EVENT(ev_begin);
    GET_EVENT(ev_end);
    CK(cudaEventRecord(ev_begin, ctx->stream));

    PropsArgs<T> pa;
    pa.nS = nE; pa.Nz = Nz; pa.nAer = nAer; pa.ldspec = (size_t)nBand;
    pa.tau_rayl = (const T*)d->tau_rayl + e0;
    pa.tau_abs = (const T*)d->tau_abs + e0;
    pa.varpi_cab = d->varpi_cabannes;
    pa.aer_tau_mod = cT + o_taumod; pa.aer_varpi_mod = cT + o_varmod;
    pa.rayl_allzero = cI + o_rz;
    pa.tau = pw; pa.varpi = pw + (size_t)Nz * nE; pa.tausum = pw + (size_t)2 * Nz * nE;
    pa.W = pw + (size_t)(2 * Nz + Nz + 1) * nE;
    k_props<T><<<(nE + 127) / 128, 128, 0, ctx->stream>>>(pa);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(rrs::launch_fscatt(ctx->stream, nE, Nz, nAer, (size_t)nBand, (const T*)d->tau_rayl + e0, cT + o_taumod, fsc));
    ctx->launches++;

    rrs::Args a;
    memset(&a, 0, sizeof a);
    a.N = N; a.nStokes = nSt; a.Nz = Nz; a.nC = nC; a.nVZA = nV; a.nM = nM;
    a.nBand = nBand; a.e0 = e0; a.nE = nE; a.n1lo = n1lo; a.n1cnt = n1cnt; a.nRaman = nR;
    a.istart = (d->imu0 - 1) * nSt;
    a.order = ctx->rrs_order;
    a.neumann = ctx->inverse_mode;
    a.u_lo = e0; a.u_n0cnt = nE;
    a.mu0 = d->mu0;
    a.mu = cT + o_mu; a.wt = cT + o_wt; a.sgn = cT + o_sg;
    a.shift = cI + o_sh; a.wR = cT + o_wR; a.ndoubl = cI + o_nd;
    a.p_tau = pa.tau; a.p_varpi = pa.varpi; a.p_tausum = pa.tausum; a.p_W = pa.W; a.fsc = fsc;
    a.Zpp = (const T*)d->Zpp; a.Zmp = (const T*)d->Zmp;
    a.ZppR = (const T*)rd->Zpp_raman; a.ZmpR = (const T*)rd->Zmp_raman;
    a.F0 = (const T*)d->F0;
    a.surf_rho = lambertian ? cT + o_rho : (const T*)d->surf_rho;
    a.lambertian = lambertian ? 1 : 0;
    a.rec0 = rec0; a.rec1 = rec1; a.irec0 = i0; a.irec1 = i1; a.uout = uo;
    a.vza_row = cI + o_vrow; a.vza_w = cT + o_vw;
    a.R_SFI = R_SFI; a.T_SFI = T_SFI;
    CK(set_smem(ctx, kd->k_elastic, kd->smem_elastic));
    CK(set_smem(ctx, kd->k_inel, kd->smem_inel));
    size_t total_units = U;
    if (a.order == 1) {
        // donor-major list of the VALID units of this owner range: consecutive warps share the donor-side records.
        // Built on the host once per (band, owner range, shift table) and kept on the device.
        std::vector<int> key = {nBand, n1lo, n1cnt};
        key.insert(key.end(), rd->i_shift, rd->i_shift + nR);
        if (key != ctx->rrs_key || !ctx->rrsunits.p) {
            std::vector<int2> ul;
            ul.reserve(U);
            for (int n0 = e0; n0 < e1; ++n0)
                for (int k = 0; k < nR; ++k) {
                    const int n1 = n0 - rd->i_shift[k];
                    if (n1 >= n1lo && n1 < n1lo + n1cnt) ul.push_back(make_int2(n1, k));
                }
            CK(cudaStreamSynchronize(ctx->stream));      // the previous call may still read the old list
            if ((rc = ensure(ctx, ctx->rrsunits, std::max<size_t>(16, ul.size() * sizeof(int2))))) return rc;
            if (!ul.empty()) CK(cudaMemcpy(ctx->rrsunits.p, ul.data(), ul.size() * sizeof(int2), cudaMemcpyHostToDevice));
            ctx->rrs_key = key;
            ctx->rrs_nunits = ul.size();
        }
        total_units = ctx->rrs_nunits;
        a.units = (const int2*)ctx->rrsunits.p;
    }
    if (total_units > 0xfffffff0u) return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_rrs: too many units for one launch");
    int be = 0, bi = 0;
    for (int m = 0; m < nM; ++m) {
        for (int step = 0; step <= Nz; ++step) {
            const int mode = (step == Nz) ? 1 : 0;
            const int iz = mode ? Nz - 1 : step;
            a.cE_in = cE[be]; a.cE_out = cE[be ^ 1];
            a.cI_in = cIe[bi]; a.cI_out = cIe[bi ^ 1];
            GET_EVENT(x0);
            GET_EVENT(x1);
            GET_EVENT(x2);
            CK(cudaEventRecord(x0, ctx->stream));
            CK(kd->launch_elastic(ctx->stream, a, m, iz, mode));
            CK(cudaEventRecord(x1, ctx->stream));
            CK(kd->launch_inel(ctx->stream, a, m, iz, mode, (unsigned int)total_units));
            CK(cudaEventRecord(x2, ctx->stream));
            ctx->launches += 2;
            ctx->seg_add.push_back({x0, x1});
            ctx->seg_layer.push_back({x1, x2});
            be ^= 1; bi ^= 1;
        }
        CK(rrs::launch_reduce(ctx->stream, a, ieR_SFI, ieT_SFI, (m == nM - 1) ? (int*)ctx->counter.p + 16 : nullptr));
        ctx->launches++;
    }
    CK(cudaEventRecord(ev_end, ctx->stream));
    ctx->ev[0] = ev_begin;
    ctx->ev[1] = ev_end;
    ctx->timing_valid = true;
    return VSM_OK;
}

extern "C" int vsm_rt_column_rrs_f64(vsm_ctx* c, const vsm_column_desc* d, const vsm_rrs_desc* r, double* R, double* T_, double* ieR,
                                     double* ieT) {
    // The unit state is (nRaman x owner points) x 8.4 KB (NP = 16), double-buffered: when it exceeds the workspace limit the owner
    // range is cut into sub-ranges that run one after the other (each recomputes the elastic column of its own halo; results are
    // identical because a unit depends on nothing but the elastic state of n1 and n0).
    if (c && d && r && r->struct_size == (int)sizeof(vsm_rrs_desc) && r->n1_count > 0 && r->n_raman > 0 && d->N > 0 && d->nVZA > 0) {
        const rrs::Desc* kd = rrs::desc(d->N);
        if (kd) {
            if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, VSM_ERR_CUDA, "cudaSetDevice failed");
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return fail(c, VSM_ERR_CUDA, "cudaMemGetInfo failed");
            uint64_t limit = c->ws_limit ? c->ws_limit : (uint64_t)48 << 30;
            limit = std::min<uint64_t>(limit, ((uint64_t)free_b + c->rrs.cap) * 8 / 10);
            const uint64_t per_pt = (uint64_t)r->n_raman * (2 * kd->ci + 2 * (uint64_t)d->nVZA * d->nStokes) * sizeof(double);
            int ndm = 1;
            for (int iz = 0; d->ndoubl && iz < d->Nz; ++iz) ndm = std::max(ndm, d->ndoubl[iz]);
            const uint64_t elastic = (uint64_t)d->nSpec * ((uint64_t)ndm * (kd->rec0 + kd->rec1) + kd->irec0 + kd->irec1 + 2 * kd->ce + 64) * sizeof(double);
            if (elastic + per_pt * r->n1_count > limit && limit > elastic + per_pt) {
                const int chunk = (int)std::max<uint64_t>(1, (limit - elastic) / per_pt);
                const size_t nOut = (size_t)d->nVZA * d->nStokes;
                for (int lo = 0; lo < r->n1_count; lo += chunk) {
                    vsm_rrs_desc sub = *r;
                    sub.n1_begin = r->n1_begin + lo;
                    sub.n1_count = std::min(chunk, r->n1_count - lo);
                    const int rc = rt_column_rrs_impl(c, d, &sub, R + lo * nOut, T_ + lo * nOut, ieR + lo * nOut, ieT + lo * nOut);
                    if (rc) return rc;
                }
                return VSM_OK;
            }
        }
    }
    return rt_column_rrs_impl(c, d, r, R, T_, ieR, ieT);
}

// ----------------------------------------------------------------------------------------------------------------
// L2, single scattering: rt_run_ss(::noRS, model, iBand) rt_run.jl:593-755
// ----------------------------------------------------------------------------------------------------------------
template <typename T>
static int rt_column_ss_impl(vsm_ctx* ctx, const vsm_column_desc* d, T* R_SFI, T* T_SFI, T* hem_R, T* hem_T) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    if (!d || (d->struct_size != (int)sizeof(vsm_column_desc) && d->struct_size != (int)offsetof(vsm_column_desc, thermal_B)))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss: descriptor NULL or struct_size mismatch");
    if (d->struct_size == (int)sizeof(vsm_column_desc) && (d->thermal_B || d->sif0))
        return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_ss: extra sources are not part of this driver");
    if (d->nSpec == 0) return VSM_OK;
    if (!R_SFI || !T_SFI) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss: R_SFI/T_SFI NULL");
    const int N = d->N, nSt = d->nStokes, Nz = d->Nz, nAer = d->nAer, nV = d->nVZA, nM = d->n_moments, nSpec = d->nSpec;
    const int nOut = nV * nSt;
    if (N <= 0 || nSt < 1 || nSt > 4 || N % nSt != 0 || Nz <= 0 || nAer < 0 || nAer + 1 > kMaxComp || nV <= 0 || nM <= 0 ||
        nSpec < 0 || d->imu0 < 1 || d->imu0 * nSt > N)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss: bad dimensions");
    if (!d->mu || !d->wt || !d->vza_weight || !d->rayl_allzero || !d->vza_stream || !d->tau_rayl || !d->tau_abs || !d->F0 ||
        !d->Zpp || !d->Zmp || (nAer > 0 && (!d->aer_omega || !d->aer_ft || !d->tau_aer)))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss: NULL input array");
    if (d->surface_kind != VSM_SURFACE_LAMBERTIAN_SCALAR && d->surface_kind != VSM_SURFACE_BRDF_FOURIER)
        return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_ss: surface_kind must be LAMBERTIAN_SCALAR or BRDF_FOURIER");
    const void* surf_rho = d->surface_kind == VSM_SURFACE_BRDF_FOURIER ? d->surf_rho : nullptr;
    if (d->surface_kind == VSM_SURFACE_BRDF_FOURIER && !surf_rho) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss: BRDF_FOURIER needs surf_rho");
    CK(cudaSetDevice(ctx->device));
    const int nC = nAer + 1;
    std::vector<T> fT;
    std::vector<int> fI;
    auto pushT = [&](const double* p, size_t n) { size_t o = fT.size(); for (size_t i = 0; i < n; ++i) fT.push_back((T)p[i]); return o; };
    const size_t o_mu = pushT(d->mu, N), o_wt = pushT(d->wt, N);
    std::vector<double> taumod((size_t)std::max(1, nAer * Nz)), varmod((size_t)std::max(1, nAer));
    for (int c = 0; c < nAer; ++c) {
        const T ft = (T)d->aer_ft[c], om = (T)d->aer_omega[c];
        varmod[c] = (double)(T)((T(1) - ft) * om / (T(1) - ft * om));
        for (int iz = 0; iz < Nz; ++iz) taumod[c + (size_t)nAer * iz] = (double)(T)((T(1) - ft * om) * (T)d->tau_aer[c + (size_t)nAer * iz]);
    }
    const size_t o_taumod = pushT(taumod.data(), taumod.size()), o_varmod = pushT(varmod.data(), varmod.size());
    const size_t o_vw = pushT(d->vza_weight, (size_t)nSt * nV * nM);
    for (int iz = 0; iz < Nz; ++iz) fI.push_back(d->rayl_allzero[iz]);
    const size_t o_vrow = fI.size();
    for (int i = 0; i < nV; ++i) {
        if (d->vza_stream[i] < 1 || d->vza_stream[i] * nSt > N) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss: vza_stream out of range");
        fI.push_back((d->vza_stream[i] - 1) * nSt);
    }
    const size_t bytesT = fT.size() * sizeof(T), bytesI = fI.size() * sizeof(int);
    const size_t offI = (bytesT + 15) & ~size_t(15);
    int rc = ensure(ctx, ctx->consts, offI + bytesI);
    if (rc) return rc;
    if (ctx->stage_cap < offI + bytesI) {
        if (ctx->stage) { CK(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->stage); ctx->stage = nullptr; ctx->stage_cap = 0; }
        CK(cudaMallocHost(&ctx->stage, (offI + bytesI) * 2));
        ctx->stage_cap = (offI + bytesI) * 2;
    }
    if (!ctx->stage_done) CK(cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming | cudaEventBlockingSync));
    else CK(cudaEventSynchronize(ctx->stage_done));
    memcpy(ctx->stage, fT.data(), bytesT);
    memcpy((char*)ctx->stage + offI, fI.data(), bytesI);
    CK(cudaMemcpyAsync(ctx->consts.p, ctx->stage, offI + bytesI, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->stage_done, ctx->stream));
    const T* cT = (const T*)ctx->consts.p;
    const int* cI = (const int*)((char*)ctx->consts.p + offI);
    const size_t props_pt = ((size_t)Nz * (2 + nC) + (Nz + 1)) * sizeof(T);
    const size_t part_pt = (size_t)2 * nM * nOut * sizeof(T);
    if ((rc = ensure(ctx, ctx->props, props_pt * nSpec))) return rc;
    if ((rc = ensure(ctx, ctx->partial, part_pt * nSpec))) return rc;
    if ((rc = ensure(ctx, ctx->counter, 256))) return rc;
    CK(cudaMemsetAsync((int*)ctx->counter.p + 16, 0, 4, ctx->stream));
    ctx->ev_used = 0;
    ctx->seg_layer.clear();
    ctx->seg_add.clear();
    GET_EVENT(ev_begin);
    GET_EVENT(ev_end);
    CK(cudaEventRecord(ev_begin, ctx->stream));
    T* pw = (T*)ctx->props.p;
    PropsArgs<T> pa;
    pa.nS = nSpec; pa.Nz = Nz; pa.nAer = nAer; pa.ldspec = (size_t)nSpec;
    pa.tau_rayl = (const T*)d->tau_rayl; pa.tau_abs = (const T*)d->tau_abs;
    pa.varpi_cab = (T)d->varpi_cabannes;
    pa.aer_tau_mod = cT + o_taumod; pa.aer_varpi_mod = cT + o_varmod; pa.rayl_allzero = cI;
    pa.tau = pw; pa.varpi = pw + (size_t)Nz * nSpec; pa.tausum = pw + (size_t)2 * Nz * nSpec;
    pa.W = pw + (size_t)(2 * Nz + Nz + 1) * nSpec;
    k_props<T><<<(nSpec + 127) / 128, 128, 0, ctx->stream>>>(pa);
    ctx->launches++;
    CK(cudaGetLastError());
    ColArgs<T> a;
    memset(&a, 0, sizeof a);
    a.N = N; a.nStokes = nSt; a.Nz = Nz; a.nC = nC; a.nVZA = nV; a.istart = (d->imu0 - 1) * nSt; a.nS = nSpec;
    a.mu0 = (T)d->mu0; a.albedo = (T)d->albedo; a.surf_rho = (const T*)surf_rho;
    a.mu = cT + o_mu; a.wt = cT + o_wt;
    a.Zpp = (const T*)d->Zpp; a.Zmp = (const T*)d->Zmp;
    a.p_tau = pa.tau; a.p_varpi = pa.varpi; a.p_tausum = pa.tausum; a.p_W = pa.W;
    a.F0 = (const T*)d->F0; a.F0_off = 0;
    a.vza_row = cI + o_vrow; a.vza_w = cT + o_vw;
    T* part = (T*)ctx->partial.p;
    const size_t per_m = (size_t)nSpec * nOut;
    a.part_R = part; a.part_T = part + (size_t)nM * per_m;
    const unsigned total = (unsigned)((size_t)nM * nSpec);
    k_ss<T><<<(total + 3) / 4, 128, 0, ctx->stream>>>(a, total, hem_R, hem_T);
    ctx->launches++;
    CK(cudaGetLastError());
    k_reduce<T><<<(unsigned)((per_m + 255) / 256), 256, 0, ctx->stream>>>(nM, per_m, a.part_R, a.part_T, (const T*)nullptr, R_SFI, T_SFI,
                                                                         (T*)nullptr, (const T*)nullptr, pa.tausum + (size_t)Nz * nSpec,
                                                                         (T)d->mu0, nOut, (int*)ctx->counter.p + 16);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev_end, ctx->stream));
    ctx->ev[0] = ev_begin;
    ctx->ev[1] = ev_end;
    ctx->timing_valid = true;
    return VSM_OK;
}

extern "C" int vsm_rt_column_ss_f64(vsm_ctx* c, const vsm_column_desc* d, double* R, double* T_, double* hR, double* hT) { return rt_column_ss_impl<double>(c, d, R, T_, hR, hT); }
extern "C" int vsm_rt_column_ss_f32(vsm_ctx* c, const vsm_column_desc* d, float* R, float* T_, float* hR, float* hT) { return rt_column_ss_impl<float>(c, d, R, T_, hR, hT); }

// ----------------------------------------------------------------------------------------------------------------
// L2, single scattering with rotational Raman: rt_run_ss(::RRS, model, iBand) rt_run.jl:593-755
// ----------------------------------------------------------------------------------------------------------------
static int rt_column_ss_rrs_impl(vsm_ctx* ctx, const vsm_column_desc* d, const vsm_rrs_desc* rd, double* R_SFI, double* T_SFI,
                                 double* ieR_SFI, double* ieT_SFI, double* hem_R, double* hem_T) {
    using T = double;
    if (!ctx) return VSM_ERR_INVALID_ARG;
    if (!rd || rd->struct_size != (int)sizeof(vsm_rrs_desc) || rd->n_raman <= 0 || !rd->i_shift || !rd->varpi_raman || !rd->Zpp_raman ||
        !rd->Zmp_raman)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss_rrs: rrs descriptor NULL / incomplete");
    if (!ieR_SFI || !ieT_SFI) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss_rrs: output NULL");
    if (!d) return fail(ctx, VSM_ERR_INVALID_ARG, "rt_column_ss_rrs: descriptor NULL");
    if (rd->n1_begin != 0 || rd->n1_count != d->nSpec)
        return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_column_ss_rrs: the owner range must be the whole band");
    // elastic single scattering (validates the column descriptor; leaves tau / tau_sum of the band in ctx->props)
    int rc = rt_column_ss_impl<T>(ctx, d, R_SFI, T_SFI, hem_R, hem_T);
    if (rc) return rc;
    const int N = d->N, nSt = d->nStokes, Nz = d->Nz, nAer = d->nAer, nV = d->nVZA, nM = d->n_moments, nBand = d->nSpec, nR = rd->n_raman;
    if (nBand == 0) return VSM_OK;
    const size_t nOut = (size_t)nV * nSt;
    // constants of this stage + fScattRayleigh in the Raman workspace (ctx->consts is still read by the elastic kernels)
    std::vector<T> fT;
    std::vector<int> fI;
    for (int i = 0; i < N; ++i) fT.push_back(d->mu[i]);
    const size_t o_wR = fT.size();
    for (int k = 0; k < nR; ++k) fT.push_back(rd->varpi_raman[k]);
    const size_t o_tm = fT.size();
    std::vector<T> taumod((size_t)std::max(1, nAer * Nz));
    for (int c = 0; c < nAer; ++c)
        for (int iz = 0; iz < Nz; ++iz) taumod[c + (size_t)nAer * iz] = (T(1) - d->aer_ft[c] * d->aer_omega[c]) * d->tau_aer[c + (size_t)nAer * iz];
    for (T v : taumod) fT.push_back(v);
    const size_t o_vw = fT.size();
    for (size_t i = 0; i < (size_t)nSt * nV * nM; ++i) fT.push_back(d->vza_weight[i]);
    for (int k = 0; k < nR; ++k) fI.push_back(rd->i_shift[k]);
    const size_t o_vrow = fI.size();
    for (int i = 0; i < nV; ++i) fI.push_back((d->vza_stream[i] - 1) * nSt);
    const size_t bytesT = ((fT.size() * sizeof(T)) + 15) & ~size_t(15), bytesI = ((fI.size() * sizeof(int)) + 15) & ~size_t(15);
    const size_t fsc_bytes = (size_t)Nz * nBand * sizeof(T);
    if ((rc = ensure(ctx, ctx->rrs, bytesT + bytesI + fsc_bytes + 64))) return rc;
    if (ctx->stage_cap < bytesT + bytesI) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->stage) { cudaFreeHost(ctx->stage); ctx->stage = nullptr; ctx->stage_cap = 0; }
        CK(cudaMallocHost(&ctx->stage, (bytesT + bytesI) * 2));
        ctx->stage_cap = (bytesT + bytesI) * 2;
    }
    CK(cudaEventSynchronize(ctx->stage_done));               // the elastic stage's constant upload has left the pinned buffer
    memcpy(ctx->stage, fT.data(), fT.size() * sizeof(T));
    memcpy((char*)ctx->stage + bytesT, fI.data(), fI.size() * sizeof(int));
    CK(cudaMemcpyAsync(ctx->rrs.p, ctx->stage, bytesT + bytesI, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->stage_done, ctx->stream));
    const T* cT = (const T*)ctx->rrs.p;
    const int* cI = (const int*)((char*)ctx->rrs.p + bytesT);
    T* fsc = (T*)((char*)ctx->rrs.p + bytesT + bytesI);
    CK(rrs::launch_fscatt(ctx->stream, nBand, Nz, nAer, (size_t)nBand, (const T*)d->tau_rayl, cT + o_tm, fsc));
    ctx->launches++;
    const size_t out_bytes = nOut * nBand * sizeof(T);
    CK(cudaMemsetAsync(ieR_SFI, 0, out_bytes, ctx->stream));
    CK(cudaMemsetAsync(ieT_SFI, 0, out_bytes, ctx->stream));
    const T* pw = (const T*)ctx->props.p;                     // k_props layout of rt_column_ss_impl: tau | varpi | tausum | W
    rrs::SsArgs a;
    memset(&a, 0, sizeof a);
    a.N = N; a.nStokes = nSt; a.Nz = Nz; a.nVZA = nV; a.nBand = nBand; a.nRaman = nR; a.istart = (d->imu0 - 1) * nSt;
    a.mu = cT; a.shift = cI; a.wR = cT + o_wR;
    a.p_tau = pw; a.p_tausum = pw + (size_t)2 * Nz * nBand; a.fsc = fsc;
    a.ZppR = (const T*)rd->Zpp_raman; a.ZmpR = (const T*)rd->Zmp_raman;
    a.F0 = (const T*)d->F0;
    a.vza_row = cI + o_vrow; a.vza_w = cT + o_vw;
    a.ieR = ieR_SFI; a.ieT = ieT_SFI;
    for (int m = 0; m < nM; ++m) {
        CK(rrs::launch_ss_rrs(ctx->stream, a, m, (unsigned int)nBand));
        ctx->launches++;
    }
    GET_EVENT(ev_end2);
    CK(cudaEventRecord(ev_end2, ctx->stream));
    ctx->ev[1] = ev_end2;
    return VSM_OK;
}

extern "C" int vsm_rt_column_ss_rrs_f64(vsm_ctx* c, const vsm_column_desc* d, const vsm_rrs_desc* r, double* R, double* T_, double* ieR,
                                        double* ieT, double* hR, double* hT) {
    return rt_column_ss_rrs_impl(c, d, r, R, T_, ieR, ieT, hR, hT);
}

extern "C" int vsm_last_column_ms(vsm_ctx* ctx, float* total_ms, float* doubling_ms, float* adding_ms) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    if (!ctx->timing_valid) return fail(ctx, VSM_ERR_INVALID_ARG, "vsm_last_column_ms: no column call yet");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev[1]));
    float t = 0, tl = 0, ta = 0, x = 0;
    CK(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
    for (auto& s : ctx->seg_layer) { CK(cudaEventElapsedTime(&x, s.first, s.second)); tl += x; }
    for (auto& s : ctx->seg_add) { CK(cudaEventElapsedTime(&x, s.first, s.second)); ta += x; }
    if (total_ms) *total_ms = t;
    if (doubling_ms) *doubling_ms = tl;
    if (adding_ms) *adding_ms = ta;
    return VSM_OK;
}

extern "C" int vsm_last_column_status(vsm_ctx* ctx, int* nonfinite_outputs) {
    if (!ctx || !nonfinite_outputs) return VSM_ERR_INVALID_ARG;
    if (!ctx->timing_valid || !ctx->counter.p) return fail(ctx, VSM_ERR_INVALID_ARG, "vsm_last_column_status: no column call yet");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(nonfinite_outputs, (int*)ctx->counter.p + 16, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VSM_OK;
}

extern "C" int vsm_last_column_launches(vsm_ctx* ctx, int* layer_launches, int* adding_launches) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    if (!ctx->timing_valid) return fail(ctx, VSM_ERR_INVALID_ARG, "vsm_last_column_launches: no column call yet");
    if (layer_launches) *layer_launches = (int)ctx->seg_layer.size();
    if (adding_launches) *adding_launches = (int)ctx->seg_add.size();
    return VSM_OK;
}

// ---------------------------------------------------------------------------
// L1: one layer, one Fourier moment  (rt_kernel!(::noRS, ...), rt_kernel.jl:175-250)
// ---------------------------------------------------------------------------
template <typename T>
static int rt_layer_impl(vsm_ctx* ctx, const vsm_layer_desc* d) {
    if (!ctx) return VSM_ERR_INVALID_ARG;
    if (!d || d->struct_size != (int)sizeof(vsm_layer_desc))
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_layer: descriptor NULL or struct_size mismatch");
    const int N = d->N, nSt = d->nStokes, nSpec = d->nSpec;
    if (N <= 0 || nSt < 1 || nSt > 4 || N % nSt != 0 || nSpec < 0 || d->imu0 < 1 || d->imu0 * nSt > N ||
        d->ndoubl < 0 || d->ndoubl > 60 || d->iface < 0 || d->iface > 3 || d->m < 0 || d->iz < 1)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_layer: bad dimensions / control values");
    if (nSpec == 0) return VSM_OK;
    if (!d->tau || !d->varpi || !d->Zpp || !d->Zmp || !d->tau_sum || !d->F0 || !d->mu || !d->wt || !d->Rmp ||
        !d->Rpm || !d->Tpp || !d->Tmm || !d->J0p || !d->J0m)
        return fail(ctx, VSM_ERR_INVALID_ARG, "rt_layer: NULL input array");
    CK(cudaSetDevice(ctx->device));
    const KernelCfg cfg = pick_cfg(N);
    const int ld = N | 1;
    const size_t msz = (size_t)ld * N * sizeof(T);
    const size_t smax = (size_t)ctx->max_smem_optin;
    const size_t lv = layer_vec_bytes<T>(N), av = add_vec_bytes<T>(N);
    if (lv + 1024 > smax) return fail(ctx, VSM_ERR_UNSUPPORTED, "rt_layer: N too large");
    const int l_mats = (int)std::min<size_t>(kLayerMats, (smax - lv) / msz);
    const int a_mats = (int)std::min<size_t>(kAddMats, (smax - av) / msz);
    const size_t l_smem = lv + (size_t)l_mats * msz, a_smem = av + (size_t)a_mats * msz;
    const size_t l_slab = (size_t)(kLayerMats - l_mats) * msz, a_slab = (size_t)(kAddMats - a_mats) * msz;
    const void* kl = nullptr; const void* ka = nullptr; const void* ks = nullptr;
    // N = 45: the specialised tensor-core kernels of the column path (everything in shared memory)
    const bool fast = has_fast(N) && l_mats == kLayerMats && a_mats == kAddMats;
    const int l_threads = fast ? fast_cfg<T, 45>().layer_threads : cfg.threads;
    get_kernel_ptrs<T>(cfg.tm, fast, N, &kl, &ka);
    if (fast) ks = step_fast_kernel_ptr<T, 45>();
    else switch (cfg.tm) {
        case 1: ks = step_kernel_ptr<T, 1>(); break;
        case 2: ks = step_kernel_ptr<T, 2>(); break;
        case 3: ks = step_kernel_ptr<T, 3>(); break;
        default: ks = step_kernel_ptr<T, 4>(); break;
    }
    CK(set_smem(ctx, (const void*)kl, l_smem));
    CK(set_smem(ctx, (const void*)ks, a_smem));
    const int l_grid_max = ctx->num_sms * occupancy<T>(kl, l_threads, l_smem);
    const int s_grid_max = ctx->num_sms * occupancy<T>(ks, l_threads, a_smem);
    const size_t NN2 = (size_t)N * N;
    // AddedLayer storage: the caller's arrays when given, library scratch otherwise
    const bool own_added = !(d->rmp && d->tpp && d->j0p && d->j0m);
    int rc;
    if (own_added && (rc = ensure(ctx, ctx->added, ((size_t)2 * NN2 + 2 * N) * nSpec * sizeof(T)))) return rc;
    T* ad = (T*)ctx->added.p;
    T* add_r = own_added ? ad : (T*)d->rmp;
    T* add_t = own_added ? ad + NN2 * nSpec : (T*)d->tpp;
    T* add_jp = own_added ? ad + 2 * NN2 * nSpec : (T*)d->j0p;
    T* add_jm = own_added ? add_jp + (size_t)N * nSpec : (T*)d->j0m;
    if ((rc = ensure(ctx, ctx->slabs, std::max<size_t>((size_t)l_slab * l_grid_max + (size_t)a_slab * s_grid_max, 16)))) return rc;
    if ((rc = ensure(ctx, ctx->counter, 256))) return rc;
    const int ctl[4] = {d->ndoubl, 1, d->iface, 0};    // ndoubl[0], scatter[0], iface[0], layer_order[0]
    if ((rc = ensure(ctx, ctx->consts, 64))) return rc;
    CK(cudaMemcpyAsync(ctx->consts.p, ctl, sizeof ctl, cudaMemcpyHostToDevice, ctx->stream));
    const int* cI = (const int*)ctx->consts.p;
    if (d->scatter) {
        if ((rc = ensure(ctx, ctx->geom, ((size_t)2 * N * N + 2 * N) * sizeof(T) + (size_t)N * N + N))) return rc;
        T* geo1 = (T*)ctx->geom.p; T* geo2 = geo1 + (size_t)N * N;
        T* gs1 = geo2 + (size_t)N * N; T* gs2 = gs1 + N;
        unsigned char* gcls = (unsigned char*)(gs2 + N);
        k_geom<T><<<(N * N + 255) / 256, 256, 0, ctx->stream>>>(N, nSt, (d->imu0 - 1) * nSt, (const T*)d->mu,
                                                                (const T*)d->wt, geo1, geo2, gs1, gs2, gcls);
        ctx->launches++;
        CK(cudaGetLastError());
        ColArgs<T> a;
        memset(&a, 0, sizeof a);
        a.N = N; a.nStokes = nSt; a.Nz = 1; a.nC = 1; a.nVZA = 0; a.ld = ld;
        a.istart = (d->imu0 - 1) * nSt; a.nS = nSpec; a.m0 = d->m; a.Mb = 1; a.n_sched_layers = 1;
        a.n_smem_mats = l_mats; a.z_spectral = d->z_is_spectral ? 1 : 0; a.mz0 = d->m; a.neumann = ctx->inverse_mode;
        a.mu0 = (T)d->mu0; a.albedo = T(0);
        a.mu = (const T*)d->mu; a.wt = (const T*)d->wt;
        a.ndoubl = cI; a.scatter = cI + 1; a.iface = cI + 2; a.layer_order = cI + 3; a.sched_idx = cI + 3;
        a.Zpp = (const T*)d->Zpp; a.Zmp = (const T*)d->Zmp;
        a.p_tau = (const T*)d->tau; a.p_varpi = (const T*)d->varpi; a.p_tausum = (const T*)d->tau_sum;
        a.p_W = (const T*)d->tau;                      // nC == 1: weights unused
        a.F0 = (const T*)d->F0; a.F0_off = 0;
        a.geo1 = geo1; a.geo2 = geo2; a.gs1 = gs1; a.gs2 = gs2; a.gcls = gcls;
        a.add_r = add_r; a.add_t = add_t; a.add_jp = add_jp; a.add_jm = add_jm;
        a.slab = (T*)ctx->slabs.p; a.slab_stride = l_slab / sizeof(T);
        a.counter = (unsigned int*)ctx->counter.p;
        CK(cudaMemsetAsync(ctx->counter.p, 0, 4, ctx->stream));
        const int grid = (int)std::min<size_t>((size_t)nSpec, (size_t)l_grid_max);
        if ((rc = do_launch<T>(ctx, true, cfg.tm, fast, N, a, (unsigned)nSpec, grid, l_threads, l_smem))) return rc;
    }
    {
        StepArgs<T> s;
        memset(&s, 0, sizeof s);
        s.N = N; s.nStokes = nSt; s.nS = nSpec; s.ld = ld; s.n_smem_mats = a_mats;
        s.iface = d->iface; s.first = (d->iz == 1) ? 1 : 0; s.scatter = d->scatter ? 1 : 0; s.neumann = ctx->inverse_mode;
        s.tau = (const T*)d->tau; s.mu = (const T*)d->mu;
        s.add_r = add_r; s.add_t = add_t; s.add_jp = add_jp; s.add_jm = add_jm;
        s.Rmp = (T*)d->Rmp; s.Rpm = (T*)d->Rpm; s.Tpp = (T*)d->Tpp; s.Tmm = (T*)d->Tmm;
        s.J0p = (T*)d->J0p; s.J0m = (T*)d->J0m;
        s.out_rpm = (T*)d->rpm; s.out_tmm = (T*)d->tmm;
        s.out_rmp = (T*)d->rmp; s.out_tpp = (T*)d->tpp; s.out_jp = (T*)d->j0p; s.out_jm = (T*)d->j0m;
        s.slab = (T*)ctx->slabs.p + ((size_t)l_slab * l_grid_max) / sizeof(T); s.slab_stride = a_slab / sizeof(T);
        s.counter = (unsigned int*)ctx->counter.p;
        CK(cudaMemsetAsync(ctx->counter.p, 0, 4, ctx->stream));
        const int grid = (int)std::min<size_t>((size_t)nSpec, (size_t)s_grid_max);
        cudaError_t e;
        if (fast) e = launch_step_fast<T, 45>(ctx->stream, s, (unsigned)nSpec, grid, a_smem);
        else switch (cfg.tm) {
            case 1: e = launch_step<T, 1>(ctx->stream, s, (unsigned)nSpec, grid, cfg.threads, a_smem); break;
            case 2: e = launch_step<T, 2>(ctx->stream, s, (unsigned)nSpec, grid, cfg.threads, a_smem); break;
            case 3: e = launch_step<T, 3>(ctx->stream, s, (unsigned)nSpec, grid, cfg.threads, a_smem); break;
            default: e = launch_step<T, 4>(ctx->stream, s, (unsigned)nSpec, grid, cfg.threads, a_smem); break;
        }
        ctx->launches++;
        if (e != cudaSuccess) return fail(ctx, VSM_ERR_CUDA, std::string("k_interact_step launch: ") + cudaGetErrorString(e));
    }
    return VSM_OK;
}

extern "C" int vsm_rt_layer_f64(vsm_ctx* c, const vsm_layer_desc* d) { return rt_layer_impl<double>(c, d); }
extern "C" int vsm_rt_layer_f32(vsm_ctx* c, const vsm_layer_desc* d) { return rt_layer_impl<float>(c, d); }
