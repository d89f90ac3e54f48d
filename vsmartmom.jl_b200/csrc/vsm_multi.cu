// vsm_multi.cu -- several GPUs driven from ONE host process behind the C-ABI: the Julia side
// runs a single process, so the spectral sharding lives in the library: one vsm_ctx (own stream) per device, the
// band cut into contiguous shards, every device runs vsm_rt_column_* on its shard with no data-path collective
// (the layer control flow in the descriptor is already reduced over the FULL band), and ONE ncclAllGather of the
// packed (R_SFI | T_SFI | hdr | bhr_uw | bhr_dw) rows over NVLink closes the call (ncclCommInitAll communicator,
// one rank per device).  NCCL is resolved with dlopen at vsm_multi_create so that libvsmom_b200.so itself has no
// link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/vsmom_b200.h"

namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        const char* names[] = {getenv("VSMOM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return false; }
#define VSM_SYM(f) *(void**)(&f) = dlsym(h, "nccl" #f); if (!f) { err = "NCCL symbol nccl" #f " missing"; return false; }
        VSM_SYM(CommInitAll) VSM_SYM(CommDestroy) VSM_SYM(AllGather) VSM_SYM(GroupStart) VSM_SYM(GroupEnd) VSM_SYM(GetErrorString)
#undef VSM_SYM
        return true;
    }
};
NcclApi g_nccl;
thread_local std::string g_multi_create_error;

struct DBuf { void* p = nullptr; size_t cap = 0; };
int grow(DBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    if (cudaMalloc(&b.p, bytes + bytes / 8 + 256) != cudaSuccess) { cudaGetLastError(); return 1; }
    b.cap = bytes + bytes / 8 + 256;
    return 0;
}
}  // namespace

struct vsm_multi {
    int ndev = 0;
    std::vector<int> dev;
    std::vector<vsm_ctx*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<ncclComm_t> comm;
    // per-device grow-only buffers: inputs (tau_rayl, tau_abs, F0, Zpp, Zmp, surf_rho), outputs, packed rows, gathered rows
    std::vector<DBuf> tr, ta, f0, zpp, zmp, rho, out, pack, gath, thb, sif, zrp, zrm;
    std::string err;
};

static int mfail(vsm_multi* m, int code, const std::string& s) {
    if (m) m->err = s; else g_multi_create_error = s;
    return code;
}
#define MCK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return mfail(m, VSM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define NCK(call)                                                                                      \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess) return mfail(m, VSM_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

extern "C" int vsm_multi_create(vsm_multi** out, const int* devices, int ndev) {
    if (!out || !devices || ndev < 1) return mfail(nullptr, VSM_ERR_INVALID_ARG, "vsm_multi_create: bad arguments");
    *out = nullptr;
    std::string e;
    if (!g_nccl.load(e)) return mfail(nullptr, VSM_ERR_UNSUPPORTED, "vsm_multi_create: " + e);
    vsm_multi* m = new vsm_multi();
    m->ndev = ndev;
    m->dev.assign(devices, devices + ndev);
    m->ctx.assign(ndev, nullptr); m->stream.assign(ndev, nullptr); m->comm.assign(ndev, nullptr);
    m->tr.resize(ndev); m->ta.resize(ndev); m->f0.resize(ndev); m->zpp.resize(ndev); m->zmp.resize(ndev);
    m->rho.resize(ndev); m->out.resize(ndev); m->pack.resize(ndev); m->gath.resize(ndev); m->thb.resize(ndev); m->sif.resize(ndev); m->zrp.resize(ndev); m->zrm.resize(ndev);
    for (int r = 0; r < ndev; ++r) {
        if (cudaSetDevice(devices[r]) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            vsm_multi_destroy(m);
            return mfail(nullptr, VSM_ERR_CUDA, "vsm_multi_create: cannot open a stream on device " + std::to_string(devices[r]));
        }
        if (vsm_create(&m->ctx[r], devices[r], m->stream[r]) != VSM_OK) {
            const std::string why = vsm_last_error(nullptr);
            vsm_multi_destroy(m);
            return mfail(nullptr, VSM_ERR_CUDA, "vsm_multi_create: " + why);
        }
    }
    const ncclResult_t nr = g_nccl.CommInitAll(m->comm.data(), ndev, devices);
    if (nr != ncclSuccess) {
        const std::string why = g_nccl.GetErrorString(nr);
        for (auto& c : m->comm) c = nullptr;
        vsm_multi_destroy(m);
        return mfail(nullptr, VSM_ERR_CUDA, "vsm_multi_create: ncclCommInitAll: " + why);
    }
    *out = m;
    return VSM_OK;
}

extern "C" int vsm_multi_destroy(vsm_multi* m) {
    if (!m) return VSM_OK;
    for (int r = 0; r < m->ndev; ++r) {
        cudaSetDevice(m->dev[r]);
        if (m->stream[r]) cudaStreamSynchronize(m->stream[r]);
        if (m->comm[r] && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm[r]);
        if (m->ctx[r]) vsm_destroy(m->ctx[r]);
        for (DBuf* b : {&m->tr[r], &m->ta[r], &m->f0[r], &m->zpp[r], &m->zmp[r], &m->rho[r], &m->out[r], &m->pack[r], &m->gath[r], &m->thb[r], &m->sif[r], &m->zrp[r], &m->zrm[r]})
            if (b->p) cudaFree(b->p);
        if (m->stream[r]) cudaStreamDestroy(m->stream[r]);
    }
    delete m;
    return VSM_OK;
}

extern "C" const char* vsm_multi_last_error(vsm_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
extern "C" int vsm_multi_ndev(vsm_multi* m) { return m ? m->ndev : 0; }
extern "C" vsm_ctx* vsm_multi_ctx(vsm_multi* m, int r) { return (m && r >= 0 && r < m->ndev) ? m->ctx[r] : nullptr; }

// contiguous [lo, hi) per rank; interior boundaries are multiples of 32 (vsmom_b200/sharding.py::shard_bounds)
static void shard_bounds(int nSpec, int world, int r, int* lo, int* hi, int* per_out) {
    int per = (nSpec + world - 1) / world;
    per = (per + 31) / 32 * 32;
    *lo = std::min(nSpec, r * per);
    *hi = std::min(nSpec, *lo + per);
    if (per_out) *per_out = per;
}

template <typename T>
static int column_sharded(vsm_multi* m, const vsm_column_desc* d, T* R, T* Tt, T* H, T* uw, T* dw,
                          int (*column)(vsm_ctx*, const vsm_column_desc*, T*, T*, T*, T*, T*), ncclDataType_t nt) {
    if (!m) return VSM_ERR_INVALID_ARG;
    if (!d || d->struct_size != (int)sizeof(vsm_column_desc)) return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_sharded: descriptor NULL or struct_size mismatch");
    if (!R || !Tt) return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_sharded: R_SFI/T_SFI NULL");
    if (!d->tau_rayl || !d->tau_abs || !d->F0 || !d->Zpp || !d->Zmp) return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_sharded: NULL input array");
    const int nSpec = d->nSpec, Nz = d->Nz, nSt = d->nStokes, nV = d->nVZA, N = d->N, nM = d->n_moments, nC = d->nAer + 1;
    if (nSpec <= 0) return VSM_OK;
    const int nOut = nV * nSt, W = 3 * nOut + 2 * nSt;          // packed row: R | T | hdr | bhr_uw | bhr_dw
    const int world = m->ndev;
    int per = 0;
    { int a, b; shard_bounds(nSpec, world, 0, &a, &b, &per); }
    const size_t zbytes = (size_t)N * N * nC * nM * sizeof(T);
    const bool brdf = d->surface_kind == VSM_SURFACE_BRDF_FOURIER && d->surf_rho;
    // ---- enqueue every device's shard (host calls return after enqueueing: the devices run concurrently)
    for (int r = 0; r < world; ++r) {
        int lo, hi;
        shard_bounds(nSpec, world, r, &lo, &hi, nullptr);
        const int nS = hi - lo;
        MCK(cudaSetDevice(m->dev[r]));
        cudaStream_t s = m->stream[r];
        if (grow(m->pack[r], (size_t)per * W * sizeof(T)) || grow(m->gath[r], (size_t)world * per * W * sizeof(T)))
            return mfail(m, VSM_ERR_NOMEM, "rt_column_sharded: cudaMalloc failed");
        MCK(cudaMemsetAsync(m->pack[r].p, 0, (size_t)per * W * sizeof(T), s));
        if (nS == 0) continue;
        if (grow(m->tr[r], (size_t)nS * Nz * sizeof(T)) || grow(m->ta[r], (size_t)nS * Nz * sizeof(T)) ||
            grow(m->f0[r], (size_t)nS * nSt * sizeof(T)) || grow(m->zpp[r], zbytes) || grow(m->zmp[r], zbytes) ||
            grow(m->out[r], (size_t)nS * W * sizeof(T)) || (brdf && grow(m->rho[r], (size_t)N * N * nM * sizeof(T))))
            return mfail(m, VSM_ERR_NOMEM, "rt_column_sharded: cudaMalloc failed");
        // (nSpec, Nz) column-major host arrays: rows lo..hi of every layer column
        MCK(cudaMemcpy2DAsync(m->tr[r].p, (size_t)nS * sizeof(T), (const T*)d->tau_rayl + lo, (size_t)nSpec * sizeof(T),
                              (size_t)nS * sizeof(T), Nz, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpy2DAsync(m->ta[r].p, (size_t)nS * sizeof(T), (const T*)d->tau_abs + lo, (size_t)nSpec * sizeof(T),
                              (size_t)nS * sizeof(T), Nz, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->f0[r].p, (const T*)d->F0 + (size_t)lo * nSt, (size_t)nS * nSt * sizeof(T), cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->zpp[r].p, d->Zpp, zbytes, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->zmp[r].p, d->Zmp, zbytes, cudaMemcpyHostToDevice, s));
        if (brdf) MCK(cudaMemcpyAsync(m->rho[r].p, d->surf_rho, (size_t)N * N * nM * sizeof(T), cudaMemcpyHostToDevice, s));
        vsm_column_desc dd = *d;
        dd.nSpec = nS;
        if (d->thermal_B) {      // (Nz, nSpec): the shard's columns are contiguous
            if (grow(m->thb[r], (size_t)nS * Nz * sizeof(T))) return mfail(m, VSM_ERR_NOMEM, "rt_column_sharded: cudaMalloc failed");
            MCK(cudaMemcpyAsync(m->thb[r].p, (const T*)d->thermal_B + (size_t)lo * Nz, (size_t)nS * Nz * sizeof(T), cudaMemcpyHostToDevice, s));
            dd.thermal_B = m->thb[r].p;
        }
        if (d->sif0) {           // (nStokes, nSpec)
            if (grow(m->sif[r], (size_t)nS * nSt * sizeof(T))) return mfail(m, VSM_ERR_NOMEM, "rt_column_sharded: cudaMalloc failed");
            MCK(cudaMemcpyAsync(m->sif[r].p, (const T*)d->sif0 + (size_t)lo * nSt, (size_t)nS * nSt * sizeof(T), cudaMemcpyHostToDevice, s));
            dd.sif0 = m->sif[r].p;
        }
        dd.tau_rayl = m->tr[r].p; dd.tau_abs = m->ta[r].p; dd.F0 = m->f0[r].p; dd.Zpp = m->zpp[r].p; dd.Zmp = m->zmp[r].p;
        dd.surf_rho = brdf ? m->rho[r].p : nullptr;
        T* o = (T*)m->out[r].p;
        T* oR = o; T* oT = oR + (size_t)nS * nOut; T* oH = oT + (size_t)nS * nOut; T* oU = oH + (size_t)nS * nOut; T* oD = oU + (size_t)nS * nSt;
        if (column(m->ctx[r], &dd, oR, oT, oH, oU, oD) != VSM_OK)
            return mfail(m, VSM_ERR_CUDA, std::string("rt_column_sharded (device ") + std::to_string(m->dev[r]) + "): " + vsm_last_error(m->ctx[r]));
        // pack rows [R | T | hdr | uw | dw] of spectral point s: 5 strided device copies
        T* pk = (T*)m->pack[r].p;
        MCK(cudaMemcpy2DAsync(pk, (size_t)W * sizeof(T), oR, (size_t)nOut * sizeof(T), (size_t)nOut * sizeof(T), nS, cudaMemcpyDeviceToDevice, s));
        MCK(cudaMemcpy2DAsync(pk + nOut, (size_t)W * sizeof(T), oT, (size_t)nOut * sizeof(T), (size_t)nOut * sizeof(T), nS, cudaMemcpyDeviceToDevice, s));
        MCK(cudaMemcpy2DAsync(pk + 2 * nOut, (size_t)W * sizeof(T), oH, (size_t)nOut * sizeof(T), (size_t)nOut * sizeof(T), nS, cudaMemcpyDeviceToDevice, s));
        MCK(cudaMemcpy2DAsync(pk + 3 * nOut, (size_t)W * sizeof(T), oU, (size_t)nSt * sizeof(T), (size_t)nSt * sizeof(T), nS, cudaMemcpyDeviceToDevice, s));
        MCK(cudaMemcpy2DAsync(pk + 3 * nOut + nSt, (size_t)W * sizeof(T), oD, (size_t)nSt * sizeof(T), (size_t)nSt * sizeof(T), nS, cudaMemcpyDeviceToDevice, s));
    }
    // ---- the single gather
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < world; ++r)
        NCK(g_nccl.AllGather(m->pack[r].p, m->gath[r].p, (size_t)per * W, nt, m->comm[r], m->stream[r]));
    NCK(g_nccl.GroupEnd());
    // ---- device 0 -> the caller's host arrays
    MCK(cudaSetDevice(m->dev[0]));
    std::vector<T> host((size_t)world * per * W);
    MCK(cudaMemcpyAsync(host.data(), m->gath[0].p, host.size() * sizeof(T), cudaMemcpyDeviceToHost, m->stream[0]));
    for (int r = 0; r < world; ++r) { MCK(cudaSetDevice(m->dev[r])); MCK(cudaStreamSynchronize(m->stream[r])); }
    for (int r = 0; r < world; ++r) {
        int lo, hi;
        shard_bounds(nSpec, world, r, &lo, &hi, nullptr);
        for (int sidx = lo; sidx < hi; ++sidx) {
            const T* row = host.data() + ((size_t)r * per + (sidx - lo)) * W;
            memcpy(R + (size_t)sidx * nOut, row, (size_t)nOut * sizeof(T));
            memcpy(Tt + (size_t)sidx * nOut, row + nOut, (size_t)nOut * sizeof(T));
            if (H) memcpy(H + (size_t)sidx * nOut, row + 2 * nOut, (size_t)nOut * sizeof(T));
            if (uw) memcpy(uw + (size_t)sidx * nSt, row + 3 * nOut, (size_t)nSt * sizeof(T));
            if (dw) memcpy(dw + (size_t)sidx * nSt, row + 3 * nOut + nSt, (size_t)nSt * sizeof(T));
        }
    }
    return VSM_OK;
}

extern "C" int vsm_rt_column_sharded_f64(vsm_multi* m, const vsm_column_desc* d, double* R, double* T_, double* hdr, double* uw, double* dw) {
    return column_sharded<double>(m, d, R, T_, hdr, uw, dw, vsm_rt_column_f64, ncclFloat64);
}
extern "C" int vsm_rt_column_sharded_f32(vsm_multi* m, const vsm_column_desc* d, float* R, float* T_, float* hdr, float* uw, float* dw) {
    return column_sharded<float>(m, d, R, T_, hdr, uw, dw, vsm_rt_column_f32, ncclFloat32);
}

// Rotational Raman from one process on several devices: every device gets the WHOLE band (the donor wavelength of a
// coupling may lie anywhere in it) and one contiguous OWNER range of the outputs; it recomputes the elastic column of the
// range's halo locally (vsm_rt_column_rrs_f64), so the data path has no collective; ONE ncclAllGather of the packed
// (R | T | ieR | ieT) rows closes the call.  Host pointers in (column arrays, Z tables, Raman Z tables), host outputs.
// owner ranges with equal numbers of (n1, dn) units (band-edge wavelengths have fewer in-band donors) -- the same rule as
// vsmom_b200/raman.py::rrs_shard_bounds: weight(n1) = 2 + #{dn : 0 <= n1 + shift[dn] < nSpec}
static std::vector<int> rrs_cuts(int nSpec, const vsm_rrs_desc* rd, int world) {
    std::vector<double> c((size_t)nSpec + 1, 0.0);
    for (int n1 = 0; n1 < nSpec; ++n1) {
        double w = 2.0;
        for (int k = 0; k < rd->n_raman; ++k) { const int n0 = n1 + rd->i_shift[k]; w += (n0 >= 0 && n0 < nSpec) ? 1.0 : 0.0; }
        c[n1 + 1] = c[n1] + w;
    }
    std::vector<int> cuts(world + 1, nSpec);
    cuts[0] = 0;
    for (int r = 1; r < world; ++r) {
        const double target = c[nSpec] * r / world;
        cuts[r] = (int)(std::lower_bound(c.begin(), c.end(), target) - c.begin());
        cuts[r] = std::min(std::max(cuts[r], cuts[r - 1]), nSpec);
    }
    return cuts;
}

extern "C" int vsm_rt_column_rrs_sharded_f64(vsm_multi* m, const vsm_column_desc* d, const vsm_rrs_desc* rd, double* R, double* Tt,
                                             double* ieR, double* ieT) {
    using T = double;
    if (!m) return VSM_ERR_INVALID_ARG;
    if (!d || d->struct_size != (int)sizeof(vsm_column_desc) || !rd || rd->struct_size != (int)sizeof(vsm_rrs_desc))
        return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_rrs_sharded: descriptor NULL or struct_size mismatch");
    if (!R || !Tt || !ieR || !ieT) return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_rrs_sharded: output NULL");
    if (!d->tau_rayl || !d->tau_abs || !d->F0 || !d->Zpp || !d->Zmp || !rd->Zpp_raman || !rd->Zmp_raman)
        return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_rrs_sharded: NULL input array");
    const int nSpec = d->nSpec, Nz = d->Nz, nSt = d->nStokes, nV = d->nVZA, N = d->N, nM = d->n_moments, nC = d->nAer + 1;
    if (nSpec <= 0) return VSM_OK;
    const int nOut = nV * nSt, W = 4 * nOut;
    const int world = m->ndev;
    if (!rd->i_shift || rd->n_raman <= 0) return mfail(m, VSM_ERR_INVALID_ARG, "rt_column_rrs_sharded: shift table missing");
    const std::vector<int> cuts = rrs_cuts(nSpec, rd, world);
    int per = 1;
    for (int r = 0; r < world; ++r) per = std::max(per, cuts[r + 1] - cuts[r]);
    const size_t zbytes = (size_t)N * N * nC * nM * sizeof(T), zrbytes = (size_t)N * N * nM * sizeof(T);
    const size_t band = (size_t)nSpec * Nz * sizeof(T);
    const bool brdf = d->surface_kind == VSM_SURFACE_BRDF_FOURIER && d->surf_rho;
    for (int r = 0; r < world; ++r) {
        const int lo = cuts[r], hi = cuts[r + 1];
        const int nS = hi - lo;
        MCK(cudaSetDevice(m->dev[r]));
        cudaStream_t s = m->stream[r];
        if (grow(m->pack[r], (size_t)per * W * sizeof(T)) || grow(m->gath[r], (size_t)world * per * W * sizeof(T)))
            return mfail(m, VSM_ERR_NOMEM, "rt_column_rrs_sharded: cudaMalloc failed");
        MCK(cudaMemsetAsync(m->pack[r].p, 0, (size_t)per * W * sizeof(T), s));
        if (nS == 0) continue;
        if (grow(m->tr[r], band) || grow(m->ta[r], band) || grow(m->f0[r], (size_t)nSpec * nSt * sizeof(T)) || grow(m->zpp[r], zbytes) ||
            grow(m->zmp[r], zbytes) || grow(m->zrp[r], zrbytes) || grow(m->zrm[r], zrbytes) || grow(m->out[r], (size_t)nS * W * sizeof(T)) ||
            (brdf && grow(m->rho[r], (size_t)N * N * nM * sizeof(T))))
            return mfail(m, VSM_ERR_NOMEM, "rt_column_rrs_sharded: cudaMalloc failed");
        MCK(cudaMemcpyAsync(m->tr[r].p, d->tau_rayl, band, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->ta[r].p, d->tau_abs, band, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->f0[r].p, d->F0, (size_t)nSpec * nSt * sizeof(T), cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->zpp[r].p, d->Zpp, zbytes, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->zmp[r].p, d->Zmp, zbytes, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->zrp[r].p, rd->Zpp_raman, zrbytes, cudaMemcpyHostToDevice, s));
        MCK(cudaMemcpyAsync(m->zrm[r].p, rd->Zmp_raman, zrbytes, cudaMemcpyHostToDevice, s));
        if (brdf) MCK(cudaMemcpyAsync(m->rho[r].p, d->surf_rho, (size_t)N * N * nM * sizeof(T), cudaMemcpyHostToDevice, s));
        vsm_column_desc dd = *d;
        dd.tau_rayl = m->tr[r].p; dd.tau_abs = m->ta[r].p; dd.F0 = m->f0[r].p; dd.Zpp = m->zpp[r].p; dd.Zmp = m->zmp[r].p;
        dd.surf_rho = brdf ? m->rho[r].p : nullptr;
        vsm_rrs_desc rr = *rd;
        rr.n1_begin = lo; rr.n1_count = nS; rr.Zpp_raman = m->zrp[r].p; rr.Zmp_raman = m->zrm[r].p;
        T* o = (T*)m->out[r].p;
        T* o4[4] = {o, o + (size_t)nS * nOut, o + (size_t)2 * nS * nOut, o + (size_t)3 * nS * nOut};
        if (vsm_rt_column_rrs_f64(m->ctx[r], &dd, &rr, o4[0], o4[1], o4[2], o4[3]) != VSM_OK)
            return mfail(m, VSM_ERR_CUDA, std::string("rt_column_rrs_sharded (device ") + std::to_string(m->dev[r]) + "): " + vsm_last_error(m->ctx[r]));
        T* pk = (T*)m->pack[r].p;
        for (int q = 0; q < 4; ++q)
            MCK(cudaMemcpy2DAsync(pk + q * nOut, (size_t)W * sizeof(T), o4[q], (size_t)nOut * sizeof(T), (size_t)nOut * sizeof(T), nS, cudaMemcpyDeviceToDevice, s));
    }
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < world; ++r)
        NCK(g_nccl.AllGather(m->pack[r].p, m->gath[r].p, (size_t)per * W, ncclFloat64, m->comm[r], m->stream[r]));
    NCK(g_nccl.GroupEnd());
    MCK(cudaSetDevice(m->dev[0]));
    std::vector<T> host((size_t)world * per * W);
    MCK(cudaMemcpyAsync(host.data(), m->gath[0].p, host.size() * sizeof(T), cudaMemcpyDeviceToHost, m->stream[0]));
    for (int r = 0; r < world; ++r) { MCK(cudaSetDevice(m->dev[r])); MCK(cudaStreamSynchronize(m->stream[r])); }
    double* outs[4] = {R, Tt, ieR, ieT};
    for (int r = 0; r < world; ++r) {
        const int lo = cuts[r], hi = cuts[r + 1];
        for (int sidx = lo; sidx < hi; ++sidx) {
            const T* row = host.data() + ((size_t)r * per + (sidx - lo)) * W;
            for (int q = 0; q < 4; ++q) memcpy(outs[q] + (size_t)sidx * nOut, row + q * nOut, (size_t)nOut * sizeof(T));
        }
    }
    return VSM_OK;
}
