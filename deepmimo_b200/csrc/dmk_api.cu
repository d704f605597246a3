// dmk_api.cu -- extern "C" entry points of libdmk.so (see include/dmk.h).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>

#include "dmk_fd.cuh"
#include "dmk_fd_tc.cuh"
#include "dmk_fd_ws.cuh"
#include "dmk_fd_small.cuh"
#include "dmk_fd_mma.cuh"
#include "dmk_fd_rows.cuh"
#include "dmk_td.cuh"
#include "dmk_bf.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local char g_kernel[192] = "";
// Launches in flight use different slots of g_tc_ticket: the persistent tcgen05 kernel one counter out of the lower half, the warp-level
// tensor-core kernel a (next user, finished warps) pair out of the upper half -- 256 and 128 launches, both >= the 128 grids a device
// runs concurrently, and the two families never share a slot.
std::atomic<unsigned> ticket_seq{0};
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    return fail(DMK_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Host half of the prologue: everything that is a function of the parameters only is computed
// here in IEEE float64 with the same operations NumPy applies.
int build_desc(const dmk_desc* h, bool freq_domain, dmk::DevDesc& d)
{
    using namespace dmk;
    if (!h) return fail(DMK_ERR_INVALID_ARG, "desc is NULL");
    if (h->rx_filter != 0 && h->rx_filter != 1) return fail(DMK_ERR_INVALID_ARG, "ofdm.rx_filter must be 0 or 1 (got %d)", h->rx_filter);
    for (int s = 0; s < 2; ++s) {
        const int32_t* shp = s == 0 ? h->bs_shape : h->ue_shape;
        if (shp[0] < 1 || shp[1] < 1) return fail(DMK_ERR_INVALID_ARG, "antenna shape must be >= 1 (got %d x %d)", shp[0], shp[1]);
        if (h->pattern[s] != DMK_PATTERN_ISOTROPIC && h->pattern[s] != DMK_PATTERN_HALFWAVE_DIPOLE)
            return fail(DMK_ERR_UNSUPPORTED, "unknown radiation pattern id %d", h->pattern[s]);
    }
    if (h->n_cols < 1 || h->n_cols > kMaxPaths)
        return fail(DMK_ERR_INVALID_ARG, "n_cols=%d outside [1, %d]", h->n_cols, kMaxPaths);
    if (h->num_paths < 0) return fail(DMK_ERR_INVALID_ARG, "num_paths=%d < 0", h->num_paths);
    if (!(h->bandwidth > 0)) return fail(DMK_ERR_INVALID_ARG, "bandwidth must be > 0");
    if (h->n_subcarriers < 1) return fail(DMK_ERR_INVALID_ARG, "ofdm.subcarriers must be >= 1");
    if (freq_domain) {
        if (h->n_selected < 0) return fail(DMK_ERR_INVALID_ARG, "n_selected < 0");
        if (h->n_selected > 1 && !h->subcarriers && h->subc_step == 0)
            return fail(DMK_ERR_INVALID_ARG, "subcarriers is NULL and subc_step == 0");
    }
    if (h->n_times < 0 || h->n_times > DMK_MAX_TIMES) return fail(DMK_ERR_INVALID_ARG, "n_times=%d outside [0, %d]", h->n_times, DMK_MAX_TIMES);
    if (h->n_times > 0 && !h->times) return fail(DMK_ERR_INVALID_ARG, "n_times > 0 but times is NULL");

    memset(&d, 0, sizeof(d));
    d.bs0 = h->bs_shape[0]; d.bs1 = h->bs_shape[1];
    d.ue0 = h->ue_shape[0]; d.ue1 = h->ue_shape[1];
    d.Mt = d.bs0 * d.bs1; d.Mr = d.ue0 * d.ue1; d.M = d.Mt * d.Mr;
    d.P0 = h->n_cols; d.P = h->num_paths < h->n_cols ? h->num_paths : h->n_cols;
    d.N = h->n_subcarriers; d.K = h->n_selected;
    d.T = h->n_times > 0 ? h->n_times : 1; d.has_time_axis = h->n_times > 0;
    d.fov_any = h->fov_any; d.fov_side[0] = h->fov_side_enabled[0]; d.fov_side[1] = h->fov_side_enabled[1];
    d.pat[0] = h->pattern[0]; d.pat[1] = h->pattern[1];
    d.subc = h->subcarriers; d.subc_start = h->subc_start; d.subc_step = h->subc_step;
    d.times = h->times;
    d.sp[0] = h->bs_spacing; d.sp[1] = h->ue_spacing;
    const double k = M_PI / 180.0;                         // np.deg2rad (float64): x * (pi/180)
    for (int s = 0; s < 2; ++s) {
        const double* r = s == 0 ? h->bs_rot_deg : h->ue_rot_deg;
        const double rx = r[0] * k, ry = r[1] * k;
        d.sx[s] = std::sin(rx); d.cx[s] = std::cos(rx);    // geometry.py:296,:299
        d.sy[s] = std::sin(ry); d.cy[s] = std::cos(ry);    // geometry.py:295,:298
        d.rz[s] = r[2] * k;
        const double* f = s == 0 ? h->bs_fov_deg : h->ue_fov_deg;
        const double fh = f[0] * k, fv = f[1] * k;          // geometry.py:184
        d.h_lo[s] = 0 + fh / 2;                             // :187
        d.h_hi[s] = 2 * M_PI - fh / 2;
        d.v_hi[s] = M_PI / 2 + fv / 2;                      // :190
        d.v_lo[s] = M_PI / 2 - fv / 2;
    }
    // receive low-pass filter (channel.py:193-194): only the OFDM branch reads it (the TD branch never builds path_gen gains)
    d.rx_filter = (freq_domain && h->rx_filter) ? 1 : 0;
    d.lpf_log2n = -1;
    for (int b = 0; b < 31; ++b) if ((1 << b) == d.N) d.lpf_log2n = b;
    d.lpf_batch = 1;
    if (d.rx_filter) {
        long long bmax = (32LL * 1024) / (8LL * d.N);        // x[B][N] float2 within 32 KB
        if (bmax < 1) bmax = 1;
        if (bmax > 16) bmax = 16;
        d.lpf_batch = (int)bmax;
    }
    d.in_f64 = (h->flags & DMK_FLAG_F64_INPUTS) ? 1 : 0;
    d.ts_f64 = 1.0 / h->bandwidth;                          // channel.py:223
    d.ts_f32 = (float)(1.0 / h->bandwidth);                 // channel.py:223 then float32 (NEP 50 weak scalar)
    d.n_f32 = (float)d.N;
    d.inv_n = 1.0 / (double)d.N;
    return DMK_OK;
}

void bind_arrays(dmk::DevDesc& d, const float* power, const float* phase, const float* delay,
                 const float* aoa_az, const float* aoa_el, const float* aod_az, const float* aod_el,
                 const double* ue_rot, const float* doppler, int64_t n, int32_t ld)
{
    d.power = power; d.phase = phase; d.delay = delay;
    d.az[0] = aod_az; d.el[0] = aod_el; d.az[1] = aoa_az; d.el[1] = aoa_el;
    d.ue_rot = ue_rot; d.doppler = doppler;
    d.n_users = n; d.ld = ld;
}

int check_arrays(const float* a, const float* b, const float* c, const float* e, const float* f, const float* g,
                 const float* h, int64_t n, int32_t ld, int n_cols)
{
    if (n < 0) return fail(DMK_ERR_INVALID_ARG, "n_users < 0");
    if (ld < n_cols) return fail(DMK_ERR_INVALID_ARG, "ld=%d < n_cols=%d", ld, n_cols);
    if (n > 0 && (!a || !b || !c || !e || !f || !g || !h)) return fail(DMK_ERR_INVALID_ARG, "a path matrix pointer is NULL");
    return DMK_OK;
}

// Per-device state.  cudaFuncSetAttribute and the SM count are properties of a (function, device) pair, so the
// one-time opt-in to > 48 KB of dynamic shared memory is tracked per device: a process may call the library on
// cuda:0 and then on cuda:1 (compute_channels(device=...), MacroDataset over several GPUs).
constexpr int kMaxDevices = 64;
constexpr int kSmemSmall = 72 * 1024, kSmemSmall2 = 100 * 1024, kSmemMma = 100 * 1024, kSmemWs1 = 220 * 1024, kSmemWs4 = 220 * 1024, kSmemTc = 112 * 1024,
              kSmemFast = 110 * 1024, kSmemTile = 200 * 1024;
struct DeviceState { bool ready = false; int sms = 0; };
DeviceState g_dev[kMaxDevices];
std::mutex g_dev_mu;

int device_state(const DeviceState*& out)
{
    using namespace dmk;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) return fail(DMK_ERR_UNSUPPORTED, "device ordinal %d outside [0, %d)", dev, kMaxDevices);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DeviceState& s = g_dev[dev];
    if (!s.ready) {
        e = cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || s.sms <= 0) return cuda_fail(e, "cudaDeviceGetAttribute(multiProcessorCount)");
        const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small_kernel<4>, a, kSmemSmall);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small_kernel<8>, a, kSmemSmall);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small_kernel<16>, a, kSmemSmall);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small2_kernel<4, false>, a, kSmemSmall2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small2_kernel<8, false>, a, kSmemSmall2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small2_kernel<16, false>, a, kSmemSmall2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small2_kernel<4, true>, a, kSmemSmall2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small2_kernel<8, true>, a, kSmemSmall2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_small2_kernel<16, true>, a, kSmemSmall2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 0, 1, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 4, 1, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 4, 1, true>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 8, 1, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 8, 1, true>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 0, 2, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 4, 2, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 4, 2, true>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 8, 2, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<4, 8, 2, true>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<8, 0, 1, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<8, 4, 1, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<8, 4, 1, true>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<8, 8, 1, false>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_mma_kernel<8, 8, 1, true>, a, kSmemMma);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_ws_kernel<1>, a, kSmemWs1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_ws_kernel<2>, a, kSmemWs1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_ws_kernel<4>, a, kSmemWs4);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_tc_kernel, a, kSmemTc);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_fast_kernel, a, kSmemFast);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bf_fast_kernel, a, kSmemFast);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fd_tile_kernel, a, kSmemTile);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bf_kernel, a, kSmemTile);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        s.ready = true;
    }
    out = &s;
    return DMK_OK;
}

}  // namespace

extern "C" {

#ifdef DMK_TC_TRACE
int dmk_debug_tc_trace(long long* host_out, int n)
{
    return (int)cudaMemcpyFromSymbol(host_out, dmk::g_tc_trace, sizeof(long long) * n);
}
int dmk_debug_pro_trace(long long* host_out)
{
    return (int)cudaMemcpyFromSymbol(host_out, dmk::g_pro_trace, sizeof(long long) * 16);
}
#endif

const char* dmk_last_error(void) { return g_err; }
int dmk_abi_version(void) { return DMK_ABI_VERSION; }
int64_t dmk_launch_count(void) { return g_launches.load(); }
const char* dmk_last_kernel(void) { return g_kernel; }

int dmk_channels_fd(const dmk_desc* desc, const float* power_dbw, const float* phase_deg, const float* delay_s,
                    const float* aoa_az_deg, const float* aoa_el_deg, const float* aod_az_deg, const float* aod_el_deg,
                    const double* ue_rot_deg, const float* doppler_hz, int64_t n_users, int32_t ld, void* out_c64,
                    uint8_t* fov_mask, uint8_t* valid_mask, uint8_t* clip_mask, void* cuda_stream)
{
    using namespace dmk;
    DevDesc d;
    int rc = build_desc(desc, true, d);
    if (rc) return rc;
    rc = check_arrays(power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, n_users, ld, d.P0);
    if (rc) return rc;
    if (n_users == 0 || d.K == 0) return DMK_OK;
    if (!out_c64) return fail(DMK_ERR_INVALID_ARG, "out is NULL");
    if ((long long)d.K * d.T > 0x7fffffffLL / 8) return fail(DMK_ERR_INVALID_ARG, "K*T too large");
    bind_arrays(d, power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, ue_rot_deg, doppler_hz, n_users, ld);
    d.out = reinterpret_cast<float2*>(out_c64);
    d.fov_mask = fov_mask; d.valid_mask = valid_mask; d.clip_mask = clip_mask;

    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const DeviceState* dev = nullptr;
    rc = device_state(dev);
    if (rc) return rc;
    d.n_sms = dev->sms;
    const int ncols = d.K * d.T;
    // Production path: affine subcarrier selection, no time axis, tables fit in shared memory.  A single selected subcarrier
    // given only as a device list (subc_step == 0, subcarriers != NULL) is not affine from the host's point of view: it goes to
    // the generic tile kernel, which reads the list.
    if (d.K == 1 && d.subc_step == 0 && !d.subc) d.subc_step = 1;
    const bool affine = (d.subc_step != 0) && !d.rx_filter;      // the LPF runs in the generic tile kernel
    // A handful of selected subcarriers (K <= 8; the reference's default is ONE, channel.py:61), any list: warp per user with lanes =
    // antenna rows (dmk_fd_rows.cuh).  Every other kernel spreads the K columns over lanes or MMA columns and idles there.
    if (!d.has_time_axis && !d.rx_filter && d.K <= 8 && (unsigned long long)d.M * (unsigned long long)(d.Mt > d.bs0 ? d.Mt : d.bs0) < 0xffffffffULL &&
        (desc->kernel_hint == DMK_KERNEL_AUTO || desc->kernel_hint == DMK_KERNEL_ROWS)) {
        RowsCfg rc;
        rc.mul_mt  = d.Mt  > 1 ? (unsigned)((0x100000000ULL + d.Mt - 1) / d.Mt) : 0u;
        rc.mul_bs0 = d.bs0 > 1 ? (unsigned)((0x100000000ULL + d.bs0 - 1) / d.bs0) : 0u;
        rc.mul_ue0 = d.ue0 > 1 ? (unsigned)((0x100000000ULL + d.ue0 - 1) / d.ue0) : 0u;
        const long long rgrid = (n_users + kRowsWarps - 1) / kRowsWarps;
        if (rgrid > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
        const dim3 gr((unsigned)rgrid), bl(kRowsWarps * 32);
        if (d.K <= 1)      fd_rows_kernel<1><<<gr, bl, 0, st>>>(d, rc);
        else if (d.K <= 2) fd_rows_kernel<2><<<gr, bl, 0, st>>>(d, rc);
        else if (d.K <= 4) fd_rows_kernel<4><<<gr, bl, 0, st>>>(d, rc);
        else               fd_rows_kernel<8><<<gr, bl, 0, st>>>(d, rc);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "fd_rows_kernel launch");
        g_launches.fetch_add(1);
        snprintf(g_kernel, sizeof(g_kernel), "fd_rows_kernel<warp/user,lanes=rows,K<=%d> grid=%lld", d.K <= 1 ? 1 : (d.K <= 2 ? 2 : (d.K <= 4 ? 4 : 8)), rgrid);
        return DMK_OK;
    }
    FastCfg cfg;
    size_t fast_smem = 0;
    {
        const int pc = d.P > 0 ? d.P : 1;
        auto take = [&](size_t bytes) { size_t o = fast_smem; fast_smem += (bytes + 15) & ~size_t(15); return (int)o; };
        cfg.pcap = pc;
        cfg.nA = (d.K + 15) / 16;
        cfg.off_W  = take((size_t)pc * kTKW * sizeof(float2));
        cfg.off_A  = take((size_t)8 * pc * 8 * sizeof(float4));
        cfg.off_tY = take((size_t)pc * d.bs0 * sizeof(float2));
        cfg.off_tQ = take((size_t)pc * d.Mr * d.bs1 * sizeof(float2));
        cfg.off_wA = take((size_t)pc * cfg.nA * sizeof(float2));
        cfg.off_wB = take((size_t)pc * 16 * sizeof(float2));
        // q = umulhi(m, ceil(2^32/d)) == m / d exactly while m * d < 2^32; d == 1 is handled as mul = 0 (q = 0 is wrong),
        // so the multiplier is only used for d > 1 and d == 1 takes mul = 0xffffffff + special case below.
        cfg.mul_mt  = d.Mt  > 1 ? (unsigned)((0x100000000ULL + d.Mt - 1) / d.Mt) : 0u;
        cfg.mul_bs0 = d.bs0 > 1 ? (unsigned)((0x100000000ULL + d.bs0 - 1) / d.bs0) : 0u;
    }
    const bool div_ok = (unsigned long long)d.M * (unsigned long long)(d.Mt > d.bs0 ? d.Mt : d.bs0) < 0xffffffffULL;
    // Kernel choice: tensor-core (tcgen05, FP16 hi/lo split) > packed-FP32 CUDA-core > generic tile kernel.
    // dmk_desc.kernel_hint overrides it (parity tests and A/B timing use this; the Python driver maps DMK_FD_KERNEL to it).
    const int hint = desc->kernel_hint;
    const bool want_tile = hint == DMK_KERNEL_TILE;
    const bool want_ffma = hint == DMK_KERNEL_FFMA;
    TcCfg tcfg;
    size_t tc_smem = 1024;                                   // slack for the 1024-byte alignment of the operand tiles
    {
        const int pc = d.P > 0 ? d.P : 1;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
        tcfg.pcap = pc;
        tcfg.nA = (d.K + 15) / 16;
        tcfg.mtile = d.M >= 128 ? 128 : (d.M > 32 ? 64 : (d.M > 16 ? 32 : 16));   // antenna rows per tile = tcgen05 N
        tcfg.nsub = tcfg.mtile <= 64 ? 2 : 1;                // keep a pipeline stage at 64 KB of output for small arrays
        tcfg.off_A  = take((size_t)2 * tcfg.mtile * 128);    // A_hi, A_lo
        tcfg.off_B  = take((size_t)tcfg.nsub * 2 * kTcN * 128);   // B_hi, B_lo per sub-tile
        tcfg.sY = d.bs0 | 1; tcfg.sQ = (d.Mr * d.bs1) | 1; tcfg.sA = tcfg.nA | 1; tcfg.sB = 17;   // odd strides: no bank conflicts across paths
        tcfg.off_tY = take((size_t)pc * tcfg.sY * sizeof(float2));
        tcfg.off_tQ = take((size_t)pc * tcfg.sQ * sizeof(float2));
        tcfg.wa_table = d.K <= 1024;                          // larger K: coarse phasor formed on the fly from the seed tables
        tcfg.off_wA = take(tcfg.wa_table ? (size_t)pc * tcfg.sA * sizeof(float2) : 16);
        tcfg.off_wB = take((size_t)pc * tcfg.sB * sizeof(float2));
        // seed tables [pc][41]: temporaries in the (not yet used) B operand area when wA is materialised from them,
        // a region of their own (instead of wA) otherwise -- the footprint must stay under two CTAs per SM
        tcfg.off_seed = tcfg.wa_table ? tcfg.off_B : take((size_t)pc * 41 * sizeof(float2));
        tcfg.mul_mt = cfg.mul_mt; tcfg.mul_bs0 = cfg.mul_bs0;
        tc_smem += off;
    }
    // Tiny arrays (M < 64) and FoV-sparse scenarios are per-user-overhead bound: the packed-FP32 kernel is faster there
    // (profiles/README.md); DMK_FD_KERNEL=tc still forces the tensor-core kernel.
    const bool want_tc = hint == DMK_KERNEL_TC || hint == DMK_KERNEL_TC1;
    // Where the tensor-core kernels pay: at least one full stage of 128 chunks (512 B each) per user, or a panel of >= 64 elements.
    // FoV-sparse scenarios leave few paths per user (per-user overhead dominates): packed-FP32 kernel.  Measured: profiles/README.md.
    const long long n_chunks_u = (long long)d.M * ((d.K + kTcN / 2 - 1) / (kTcN / 2));
    const bool tc_shape = (d.M >= 64 || n_chunks_u >= 128) && !d.fov_any;
    // Per-user outputs of at most 128 KB go to the warp-level tensor-core kernel (dmk_fd_mma.cuh) instead: the persistent kernel is
    // bound by its helper warps there (8x8 x K=64: 1.4 against 3.5 TB/s, 16x1 x K=1024: 3.46 against 3.61; profiles/README.md).
    const bool mma_shape = affine && !d.has_time_axis && d.M <= 1024 && d.bs0 < 65536 && d.bs1 < 65536 && d.ue0 < 65536 && d.ue1 < 65536 && d.K <= 4096 &&
                           ((reinterpret_cast<uintptr_t>(out_c64) & 15) == 0);
    const bool ws_rag = d.K % (kTcN / 2) != 0;           // the persistent kernel would cut the last chunk of every row off
    const bool mma_pref = mma_shape && hint == DMK_KERNEL_AUTO && (long long)d.M * d.K * 8 <= (ws_rag ? 256 : 128) * 1024;
    // K not a multiple of 64 (12 x n resource blocks ...): the persistent kernel cuts the last chunk of every row off; taken while the
    // padding is at most half of K (it is bound by its stores, and only valid columns are stored: K = 88 on a 1024-element panel
    // 3.7 against 2.4 TB/s; tools/rag_ab.py) and above 256 KB per user; the one-CTA-per-user kernel does not have that store path
    const bool use_tc = !mma_pref && affine && !d.has_time_axis && div_ok && (!ws_rag || (2 * (((d.K + 63) / 64) * 64 - d.K) <= d.K && hint != DMK_KERNEL_TC1)) && d.K <= 4096 &&
                        !want_tile && !want_ffma && hint != DMK_KERNEL_SMALL && hint != DMK_KERNEL_SMALL1 && hint != DMK_KERNEL_MMA && (tc_shape || want_tc);
    const bool use_tc1 = use_tc && !ws_rag && tc_smem <= (size_t)kSmemTc;   // one-CTA-per-user tensor-core kernel: fallback of the persistent one
    const bool use_fast = !use_tc1 && affine && !d.has_time_axis && div_ok && fast_smem <= (size_t)kSmemFast && !want_tile;
    const int tile_w_tc1 = (kTcN / 2) * tcfg.nsub;
    const int tile_w = use_tc1 ? tile_w_tc1 : (use_fast ? kTKW : kTK);
    const int n_ct = (ncols + tile_w - 1) / tile_w;
    // Few users: split each user's column tiles over several CTAs so the grid covers >= 4 waves.
    const long long want = 4LL * 2 * dev->sms;
    long long ksplit = (want + n_users - 1) / n_users;
    if (ksplit > n_ct) ksplit = n_ct;
    if (ksplit < 1) ksplit = 1;
    const long long grid = n_users * ksplit;
    if (grid > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
    // Warp-specialised persistent tensor-core kernel in the flat-chunk formulation (dmk_fd_ws.cuh): the production path.
    // DMK_KERNEL_TC1 keeps the one-CTA-per-user kernel, which also takes the shapes whose tables do not fit next to the
    // 64 KB of operand tiles.
    const bool want_tc1 = hint == DMK_KERNEL_TC1;
    if (use_tc && !want_tc1) {
        WsCfg w;
        memset(&w, 0, sizeof(w));
        const int pc = (d.P > 0 ? d.P : 1) + 1;                  // + the zero row the builders read for padding slots
        w.S = (d.K + kTcN / 2 - 1) / (kTcN / 2);
        w.rag = ws_rag ? 1 : 0; w.row_floats = 2 * d.K; w.last_valid = 2 * d.K - (w.S - 1) * kTcN;
        const long long n_chunks = (long long)d.M * w.S;
        w.n_chunks = (int)n_chunks;
        w.n_stages = (int)((n_chunks + kTcN - 1) / kTcN);
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
        w.off_N = take((size_t)2 * kTcN * 128);
        w.off_M = take((size_t)2 * kTcN * 128);
        w.off_tab = (int)off;
        size_t toff = 0;
        auto ttake = [&](size_t bytes) { size_t o = toff; toff += (bytes + 15) & ~size_t(15); return (int)o; };
        w.sY = d.bs0 | 1; w.sQ = (d.Mr * d.bs1) | 1; w.sB = 17; w.sL = 5; w.sS = w.S | 1;
        w.off_tY = ttake((size_t)pc * w.sY * sizeof(float2));
        w.off_tQ = ttake((size_t)pc * w.sQ * sizeof(float2));
        w.off_wB = ttake((size_t)pc * w.sB * sizeof(float2));
        w.off_wL = ttake((size_t)pc * w.sL * sizeof(float2));
        w.off_wS = ttake((size_t)pc * w.sS * sizeof(float2));
        const size_t buf_bytes = ((sizeof(TcUserBuf) + 15) & ~size_t(15)) + toff;      // [user record][tables]
        w.tab_bytes = (int)buf_bytes;
        w.mul_mt = cfg.mul_mt; w.mul_bs0 = cfg.mul_bs0;
        w.mul_s = w.S > 1 ? (unsigned)((0x100000000ULL + w.S - 1) / w.S) : 0u;
        const bool chunk_div_ok = n_chunks * (long long)w.S < 0xffffffffLL && n_chunks < 0x7fffffffLL / kTcN;
        // One helper warp prepares a user in ~30-45 k cycles (float64 prologue + tables, latency-bound).  Users whose output is
        // written faster than that (< ~400 KB) make the kernel helper-bound: they get four helper warps and eight buffers in a
        // single CTA per SM (shared memory and registers allow it because only one CTA is resident).
        const size_t per_user_bytes = (size_t)d.M * d.K * sizeof(float2);
        const size_t smem1 = 1024 + off + 2 * buf_bytes, smem4 = 1024 + off + 8 * buf_bytes;
        const int hf = desc->ws_helpers;                      // 0 = by shape, 1 or 4 pins the instantiation (tests, A/B timing)
        int n_helpers = 0;
        bool one_cta = false;                                 // one helper, but the tables leave room for a single CTA per SM only
        const size_t smem2 = 1024 + off + 4 * buf_bytes;      // two helpers, four buffers: still two CTAs per SM when <= 113200
        if (hf == 2 && smem2 <= 113200) n_helpers = 2;
        else if ((per_user_bytes <= 384 * 1024 || hf == 4 || smem1 > 113200) && smem4 <= (size_t)kSmemWs4 && hf != 1) n_helpers = 4;
        else if (smem1 <= 113200) n_helpers = 1;              // + static + 1 KB reserve: two CTAs per SM
        else if (smem1 <= (size_t)kSmemWs4) { n_helpers = 1; one_cta = true; }    // wide panels (e.g. 64 x 4): still persistent
        // Few users: split each user's stages over several CTAs so the grid covers >= 4 waves.
        long long wsplit = (want + n_users - 1) / n_users;
        if (desc->ws_split > 0) wsplit = desc->ws_split;          // (-3: two buffers per helper even where three fit -- A/B timing)
        if (wsplit > w.n_stages) wsplit = w.n_stages;
        if (wsplit < 1) wsplit = 1;
        const long long items = n_users * wsplit;
        if (n_helpers && chunk_div_ok && items < 0xffffff00LL) {
            size_t ws_smem = n_helpers == 1 ? smem1 : (n_helpers == 2 ? smem2 : smem4);
            w.bufs_per_helper = 2;
            if (n_helpers == 4 && 1024 + off + 12 * buf_bytes <= (size_t)kSmemWs4 && desc->ws_split != -3) {     // room for a third buffer per helper
                w.bufs_per_helper = 3;
                ws_smem = 1024 + off + 12 * buf_bytes;
            }
            unsigned int* tickets = nullptr;
            cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void**>(&tickets), g_tc_ticket);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGetSymbolAddress(g_tc_ticket)");
            const long long resident = (((n_helpers == 1 && !one_cta) || n_helpers == 2) ? 2LL : 1LL) * dev->sms;
            const long long pgrid = items < resident ? items : resident;
            // Programmatic dependent launch: the kernel releases its dependents at once, so the next libdmk launch can fill SMs
            // as this one's persistent CTAs retire.  Unless the caller set DMK_FLAG_INDEPENDENT_LAUNCH the kernel itself
            // waits for the previous grid (griddepcontrol.wait) before touching memory: plain stream order.
            const int pdl_wait = (desc->flags & DMK_FLAG_INDEPENDENT_LAUNCH) ? 0 : 1;
            cudaLaunchConfig_t lc;
            memset(&lc, 0, sizeof(lc));
            lc.gridDim = dim3((unsigned)pgrid); lc.blockDim = dim3((9 + n_helpers) * 32); lc.dynamicSmemBytes = ws_smem; lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            lc.attrs = at; lc.numAttrs = 1;
            unsigned int* tk = tickets + (ticket_seq.fetch_add(1) % (kTcTickets / 2));
            if (n_helpers == 1)      e = cudaLaunchKernelEx(&lc, fd_ws_kernel<1>, d, w, (int)wsplit, (unsigned)items, tk, pdl_wait);
            else if (n_helpers == 2) e = cudaLaunchKernelEx(&lc, fd_ws_kernel<2>, d, w, (int)wsplit, (unsigned)items, tk, pdl_wait);
            else                     e = cudaLaunchKernelEx(&lc, fd_ws_kernel<4>, d, w, (int)wsplit, (unsigned)items, tk, pdl_wait);
            if (e != cudaSuccess) return cuda_fail(e, "fd_ws_kernel launch");
            g_launches.fetch_add(1);
            snprintf(g_kernel, sizeof(g_kernel), "fd_ws_kernel<128 chunks x 64 sc,3xf16,%d helper%s> grid=%lld items=%lld ksplit=%lld stages=%d smem=%zu",
                     n_helpers, n_helpers > 1 ? "s" : "", pgrid, items, wsplit, w.n_stages, ws_smem);
            return DMK_OK;
        }
    }
    // Small per-user outputs (M <= 1024, K <= 4096): warp-level tensor-core kernel, see dmk_fd_mma.cuh.  Default up to 128 KB per user,
    // and for every eligible shape the persistent kernel did not take (FoV-filtered scenarios, fewer than 128 chunks per user);
    // DMK_KERNEL_MMA forces it on every eligible shape, DMK_KERNEL_SMALL keeps the CUDA-core fd_small2_kernel.
    {
        const bool mma_wanted = hint == DMK_KERNEL_MMA || mma_pref || (hint == DMK_KERNEL_AUTO && (long long)d.M * d.K * 8 <= 512 * 1024);
        if (mma_shape && mma_wanted) {
            MmaCfg mc;
            memset(&mc, 0, sizeof(mc));
            // chunk width J = 32 where the chunks of an antenna row still share base phasors in blocks of >= 4 (K / 32 a multiple of 4)
            // and a user has >= 256 chunks of 16; else 16 (measured: tools/mma_sweep.py, profiles/README.md)
            // (rows of ceil(K / 32) chunks that are a multiple of 4, or padded to one at a cost of at most a third, see below)
            const int s32 = (d.K + 31) / 32;
            const bool j32_blocks = s32 % 4 == 0 || 3 * ((s32 + 3) & ~3) <= 4 * s32;
            int J = (j32_blocks && (long long)d.M * ((d.K + 15) / 16) >= 256) ? 32 : 16;
            if (desc->ws_helpers == 16 || desc->ws_helpers == 32) J = desc->ws_helpers;      // A/B timing
            const int nt = J / 4;
            mc.S = (d.K + J - 1) / J;                                // the last chunk of a row may be cut off (K % J != 0)
            mc.ragged = d.K % J != 0;
            // Rows of 3, 6, 7, 9, 10, 11 ... chunks are padded to a multiple of 4 chunks (the padding chunks are computed and never
            // stored): blocks of 4 chunks then share a base phasor, which costs less than a float64-reduced phasor for every chunk
            // (measured, tools/mma_ragged.py: pays up to a third of padding -- K = 300: 2.2 -> 3.1 TB/s -- not for 5 -> 8 chunks)
            if (mc.S % 4 != 0 && 3 * ((mc.S + 3) & ~3) <= 4 * mc.S) { mc.S = (mc.S + 3) & ~3; mc.ragged = 1; }
            mc.R = d.M * mc.S;
            mc.n_mt = (mc.R + 15) / 16;
            mc.mul_s = mc.S > 1 ? (unsigned)((0x100000000ULL + mc.S - 1) / mc.S) : 0u;
            // chunks that share a base phasor: SB consecutive segments of one antenna row, or -- rows of 1 or 2 chunks (16 / 32
            // subcarriers) -- all segments of SB / S consecutive elements along the panel's y axis
            int sb = mc.S % 8 == 0 ? 8 : (mc.S % 4 == 0 ? 4 : 0);
            mc.lg_blk_seg = sb == 8 ? 3 : 2;
            if (sb == 0 && (mc.S == 1 || mc.S == 2)) {
                if (d.bs0 % (8 / mc.S) == 0) sb = 8; else if (d.bs0 % (4 / mc.S) == 0) sb = 4;
                mc.lg_blk_seg = mc.S == 2 ? 1 : 0;
            }
            mc.G = mc.n_mt < 2 ? mc.n_mt : 2;                        // two m-tiles per group: measured best for both chunk widths
            if (desc->ws_split > 0 && desc->ws_split < 100) mc.G = desc->ws_split < mc.n_mt ? desc->ws_split : mc.n_mt;
            size_t off = 0;
            auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
            mc.off_A    = take((size_t)2 * mc.G * 8 * 256);
            mc.off_B    = take((size_t)2 * (nt / 2) * 8 * 128);
            mc.off_list = take((size_t)32);
            mc.off_meta = take((size_t)(3 * kMmWindow + 1) * sizeof(int));
            mc.warp_bytes = (int)off;
            mc.off_warps = (int)(((d.M > 256 ? (((size_t)d.M * 8 + 15) & ~size_t(15)) : (size_t)d.M * 32) + (size_t)mc.S * sizeof(double) + 15) & ~size_t(15));
            const size_t mma_smem = (size_t)mc.off_warps + off * kMmWarps;
            const bool pair = nt == 4 && mc.n_mt >= 4 && mc.G >= 2;               // two m-tiles per k-step share the B fragments (123 registers)
            // float32 inputs, no FoV filter, isotropic patterns: instantiation without the float64-input / angle / dipole code
            const bool plain = !d.in_f64 && !d.fov_any && d.pat[0] == DMK_PATTERN_ISOTROPIC && d.pat[1] == DMK_PATTERN_ISOTROPIC && sb != 0 && !mc.ragged;
            void (*kern)(DevDesc, MmaCfg, unsigned int*) = nullptr;
            if (plain) {
                if (nt == 4 && pair) kern = sb == 8 ? fd_mma_kernel<4, 8, 2, true> : fd_mma_kernel<4, 4, 2, true>;
                else if (nt == 4)    kern = sb == 8 ? fd_mma_kernel<4, 8, 1, true> : fd_mma_kernel<4, 4, 1, true>;
                else                 kern = sb == 8 ? fd_mma_kernel<8, 8, 1, true> : fd_mma_kernel<8, 4, 1, true>;
            } else {
                if (nt == 4 && pair) kern = sb == 8 ? fd_mma_kernel<4, 8, 2, false> : (sb == 4 ? fd_mma_kernel<4, 4, 2, false> : fd_mma_kernel<4, 0, 2, false>);
                else if (nt == 4)    kern = sb == 8 ? fd_mma_kernel<4, 8, 1, false> : (sb == 4 ? fd_mma_kernel<4, 4, 1, false> : fd_mma_kernel<4, 0, 1, false>);
                else                 kern = sb == 8 ? fd_mma_kernel<8, 8, 1, false> : (sb == 4 ? fd_mma_kernel<8, 4, 1, false> : fd_mma_kernel<8, 0, 1, false>);
            }
            int ctas_per_sm = 0;
            if (mma_smem <= (size_t)kSmemMma &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kMmWarps * 32, mma_smem) == cudaSuccess && ctas_per_sm > 0) {
                // persistent grid; guided draws of 8 / 4 / 2 users (see the kernel)
                const long long warps_res = (long long)ctas_per_sm * dev->sms * kMmWarps;
                long long sgrid = (n_users + 2 * kMmWarps - 1) / (2 * kMmWarps);
                if (sgrid > (long long)ctas_per_sm * dev->sms) sgrid = (long long)ctas_per_sm * dev->sms;
                const long long warps = sgrid * kMmWarps;
                mc.users_per_warp = desc->ws_split >= 100 ? desc->ws_split - 100 : 0;         // A/B timing of the draw size
                mc.draw8_above = (unsigned)(12 * warps);
                mc.draw4_above = (unsigned)(3 * warps);
                const long long upw = mc.users_per_warp;
                (void)warps_res;
                if (n_users >= 0xfffff000LL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
                unsigned int* tickets = nullptr;
                cudaError_t e0 = cudaGetSymbolAddress(reinterpret_cast<void**>(&tickets), g_tc_ticket);
                if (e0 != cudaSuccess) return cuda_fail(e0, "cudaGetSymbolAddress(g_tc_ticket)");
                kern<<<(unsigned)sgrid, kMmWarps * 32, mma_smem, st>>>(d, mc, tickets + kTcTickets / 2 + 2 * (ticket_seq.fetch_add(1) % (kTcTickets / 4)));
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return cuda_fail(e, "fd_mma_kernel launch");
                g_launches.fetch_add(1);
                snprintf(g_kernel, sizeof(g_kernel), "fd_mma_kernel<m16n8k16,3xf16,J=%d,SB=%d,MP=%d%s> grid=%lld users/warp=%lld chunks=%d G=%d smem=%zu", J, sb, pair ? 2 : 1, plain ? ",plain" : "", sgrid, upw, mc.R, mc.G, mma_smem);
                return DMK_OK;
            }
        }
    }
    // Small arrays (M <= 16): warp-level kernels, see dmk_fd_small.cuh.  Default: the densely packed fd_small2_kernel;
    // DMK_KERNEL_SMALL1 keeps the round-1 one-warp-per-user kernel for A/B timing.
    {
        const int pc = d.P > 0 ? d.P : 1;
        const int mt = d.M <= 4 ? 4 : (d.M <= 8 ? 8 : 16);
        const bool small_shape = affine && !d.has_time_axis && div_ok && d.M <= 16 && d.K <= 4096;
        const bool small_wanted = !want_tile && !want_ffma && !want_tc;      // (shapes the persistent tensor-core kernel took never get here)
        if (small_shape && small_wanted && hint != DMK_KERNEL_SMALL1) {
            Small2Cfg sc;
            memset(&sc, 0, sizeof(sc));
            if (d.K <= 64)       { sc.n0 = 8;  sc.log0 = 3; sc.n1 = (d.K + 7) / 8;   sc.n2 = 0; }
            else if (d.K <= 256) { sc.n0 = 16; sc.log0 = 4; sc.n1 = (d.K + 15) / 16; sc.n2 = 0; }
            else                 { sc.n0 = 16; sc.log0 = 4; sc.n1 = 16;              sc.n2 = (d.K + 255) / 256; }
            const int n_seed = sc.n0 + sc.n1 + sc.n2;
            sc.strideA = mt * 16 + 16;                         // +16 B: the 32 lanes of a chain round store to 32 different pool rows
            sc.strideW = n_seed | 1;
            const int per_path = sc.strideA + sc.strideW * 8;
            // 16 warps per SM (4 CTAs of 13.5 KB per warp), 20 when at most 8 rows are held in registers (5 CTAs, 10.5 KB per warp).
            // A pool of ~32-40 paths is one dense chain round per pass.
            const int budget = mt <= 8 ? 10 * 1024 + 512 : 13 * 1024 + 512;
            int cap = (budget - 256) / per_path;
            if (cap > (mt <= 8 ? 32 : 40)) cap = mt <= 8 ? 32 : 40;
            if (cap < d.P0) cap = d.P0;                         // one user always fits
            sc.window = kS2Window;
            sc.cap = cap;
            size_t off = 0;
            auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
            sc.off_A    = take((size_t)cap * sc.strideA);
            sc.off_W    = take((size_t)cap * sc.strideW * 8);
            sc.off_list = take((size_t)cap);
            sc.off_meta = take((size_t)(4 * kS2Window + 1) * sizeof(int));
            sc.warp_bytes = (int)off;
            sc.mul_mt = cfg.mul_mt; sc.mul_bs0 = cfg.mul_bs0;
            // contiguous users per warp: 16 when that still leaves >= 2 waves of CTAs, fewer for short user ranges
            long long upw = n_users / (2LL * 3 * dev->sms * kS2Warps);
            upw = upw < 4 ? 4 : (upw > 16 ? 16 : upw);
            sc.users_per_warp = (int)upw;
            const size_t small_smem = off * kS2Warps;
            const long long sgrid = (n_users + upw * kS2Warps - 1) / (upw * kS2Warps);
            if (small_smem <= (size_t)kSmemSmall2 && sgrid <= 0x7fffffffLL) {
                const bool l3 = sc.n2 > 0;
                if (mt == 4)       { if (l3) fd_small2_kernel<4, true><<<(unsigned)sgrid, kS2Warps * 32, small_smem, st>>>(d, sc);
                                     else    fd_small2_kernel<4, false><<<(unsigned)sgrid, kS2Warps * 32, small_smem, st>>>(d, sc); }
                else if (mt == 8)  { if (l3) fd_small2_kernel<8, true><<<(unsigned)sgrid, kS2Warps * 32, small_smem, st>>>(d, sc);
                                     else    fd_small2_kernel<8, false><<<(unsigned)sgrid, kS2Warps * 32, small_smem, st>>>(d, sc); }
                else               { if (l3) fd_small2_kernel<16, true><<<(unsigned)sgrid, kS2Warps * 32, small_smem, st>>>(d, sc);
                                     else    fd_small2_kernel<16, false><<<(unsigned)sgrid, kS2Warps * 32, small_smem, st>>>(d, sc); }
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return cuda_fail(e, "fd_small2_kernel launch");
                g_launches.fetch_add(1);
                snprintf(g_kernel, sizeof(g_kernel), "fd_small2_kernel<%d rows,dense pairs> grid=%lld users/warp=%lld cap=%d smem=%zu", mt, sgrid, upw, cap, small_smem);
                return DMK_OK;
            }
        }
        SmallCfg sc;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
        sc.pcap = pc;
        sc.n_hi = (((d.K + 15) / 16) + 7) / 8;
        sc.sY = d.bs0 | 1; sc.sQ = (d.Mr * d.bs1) | 1; sc.sS = (8 + sc.n_hi) | 1;
        sc.off_sh   = take(sizeof(FdShared));
        sc.off_tY   = take((size_t)pc * sc.sY * sizeof(float2));
        sc.off_tQ   = take((size_t)pc * sc.sQ * sizeof(float2));
        sc.off_wB   = take((size_t)pc * 17 * sizeof(float2));
        sc.off_seed = take((size_t)pc * sc.sS * sizeof(float2));
        sc.off_A    = take((size_t)pc * mt * sizeof(float4));
        sc.warp_bytes = (int)off;
        sc.mul_mt = cfg.mul_mt; sc.mul_bs0 = cfg.mul_bs0;
        const size_t small_smem = off * kSmallWarps;
        const bool use_small = small_shape && small_wanted && small_smem <= (size_t)kSmemSmall;
        if (use_small) {
            const long long sgrid = (n_users + kSmallWarps - 1) / kSmallWarps;
            if (sgrid > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
            cudaError_t e = cudaSuccess;
            if (mt == 4)      fd_small_kernel<4><<<(unsigned)sgrid, kSmallWarps * 32, small_smem, st>>>(d, sc);
            else if (mt == 8) fd_small_kernel<8><<<(unsigned)sgrid, kSmallWarps * 32, small_smem, st>>>(d, sc);
            else              fd_small_kernel<16><<<(unsigned)sgrid, kSmallWarps * 32, small_smem, st>>>(d, sc);
            e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "fd_small_kernel launch");
            g_launches.fetch_add(1);
            snprintf(g_kernel, sizeof(g_kernel), "fd_small_kernel<%d rows,warp/user> grid=%lld smem=%zu", mt, sgrid, small_smem);
            return DMK_OK;
        }
    }
    if (use_tc1) {
        fd_tc_kernel<<<(unsigned)grid, kTcThreads, tc_smem, st>>>(d, tcfg, (int)ksplit);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "fd_tc_kernel launch");
        g_launches.fetch_add(1);
        snprintf(g_kernel, sizeof(g_kernel), "fd_tc_kernel<%dx128,3xf16> grid=%lld ksplit=%lld smem=%zu", tcfg.mtile, grid, ksplit, tc_smem);
        return DMK_OK;
    }
    if (use_fast) {
        fd_fast_kernel<<<(unsigned)grid, kFdThreads, fast_smem, st>>>(d, cfg, (int)ksplit);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "fd_fast_kernel launch");
        g_launches.fetch_add(1);
        snprintf(g_kernel, sizeof(g_kernel), "fd_fast_kernel<64x256,ffma2> grid=%lld ksplit=%lld smem=%zu", grid, ksplit, fast_smem);
        return DMK_OK;
    }
    size_t smem = (size_t)kMaxPaths * (kTK + kTM) * sizeof(float2);
    if (d.rx_filter) {
        smem += (size_t)(1 + d.lpf_batch) * d.N * sizeof(float2);      // twiddle table + FFT batch (dmk_fd.cuh: lpf_w_tile)
        if (smem > (size_t)kSmemTile)
            return fail(DMK_ERR_UNSUPPORTED, "ofdm.rx_filter=1 supports ofdm.subcarriers <= 9728 (got %d)", d.N);
        // what is left of 200 KB caches the transformed paths of a user (np * K complex values) across its column tiles
        d.lpf_cache = (int)((kSmemTile - smem) / sizeof(float2));
        if ((long long)d.lpf_cache > (long long)kMaxPaths * d.K) d.lpf_cache = kMaxPaths * d.K;
        smem += (size_t)d.lpf_cache * sizeof(float2);
    }
    fd_tile_kernel<<<(unsigned)grid, kFdThreads, smem, st>>>(d, (int)ksplit);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fd_tile_kernel launch");
    g_launches.fetch_add(1);
    snprintf(g_kernel, sizeof(g_kernel), "fd_tile_kernel<64x128%s> grid=%lld ksplit=%lld", d.rx_filter ? (d.lpf_log2n >= 0 ? ",lpf-fft" : ",lpf-dft") : "", grid, ksplit);
    return DMK_OK;
}

int dmk_channels_td(const dmk_desc* desc, const float* power_dbw, const float* phase_deg, const float* delay_s,
                    const float* aoa_az_deg, const float* aoa_el_deg, const float* aod_az_deg, const float* aod_el_deg,
                    const double* ue_rot_deg, const float* doppler_hz, int64_t n_users, int32_t ld, void* out_c64,
                    uint8_t* fov_mask, uint8_t* valid_mask, int32_t* path_slot, void* cuda_stream)
{
    return dmk_channels_td_tau(desc, power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, ue_rot_deg,
                               doppler_hz, n_users, ld, out_c64, fov_mask, valid_mask, path_slot, nullptr, cuda_stream);
}

int dmk_channels_td_tau(const dmk_desc* desc, const float* power_dbw, const float* phase_deg, const float* delay_s,
                        const float* aoa_az_deg, const float* aoa_el_deg, const float* aod_az_deg, const float* aod_el_deg,
                        const double* ue_rot_deg, const float* doppler_hz, int64_t n_users, int32_t ld, void* out_c64,
                        uint8_t* fov_mask, uint8_t* valid_mask, int32_t* path_slot, float* tau, void* cuda_stream)
{
    using namespace dmk;
    DevDesc d;
    int rc = build_desc(desc, false, d);
    if (rc) return rc;
    rc = check_arrays(power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, n_users, ld, d.P0);
    if (rc) return rc;
    if (n_users == 0 || d.P == 0) return DMK_OK;
    if (!out_c64) return fail(DMK_ERR_INVALID_ARG, "out is NULL");
    if (n_users > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
    bind_arrays(d, power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, ue_rot_deg, doppler_hz, n_users, ld);
    d.out = reinterpret_cast<float2*>(out_c64);
    d.fov_mask = fov_mask; d.valid_mask = valid_mask; d.path_slot = path_slot; d.tau_out = tau;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const DeviceState* dev = nullptr;
    rc = device_state(dev);
    if (rc) return rc;
    d.n_sms = dev->sms;
    if (!d.has_time_axis && desc->kernel_hint != DMK_KERNEL_TILE) {
        // the reference's own time-domain mode: one warp per user, one store instruction per antenna row (dmk_td.cuh)
        const long long wgrid = (n_users + kTdwWarps - 1) / kTdwWarps;
        td_warp_kernel<<<(unsigned)wgrid, kTdwWarps * 32, 0, st>>>(d);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "td_warp_kernel launch");
        g_launches.fetch_add(1);
        snprintf(g_kernel, sizeof(g_kernel), "td_warp_kernel<warp/user> grid=%lld", wgrid);
        return DMK_OK;
    }
    td_kernel<<<(unsigned)n_users, kTdThreads, 0, st>>>(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "td_kernel launch");
    g_launches.fetch_add(1);
    snprintf(g_kernel, sizeof(g_kernel), "td_kernel grid=%lld", (long long)n_users);
    return DMK_OK;
}

int dmk_beam_amplitude_fd(const dmk_desc* desc, const float* power_dbw, const float* phase_deg, const float* delay_s,
                          const float* aoa_az_deg, const float* aoa_el_deg, const float* aod_az_deg, const float* aod_el_deg,
                          const double* ue_rot_deg, int64_t n_users, int32_t ld, const void* beams_c64, int32_t n_beams,
                          float* mean_abs, uint8_t* fov_mask, uint8_t* valid_mask, uint8_t* clip_mask, void* cuda_stream)
{
    using namespace dmk;
    DevDesc d;
    int rc = build_desc(desc, true, d);
    if (rc) return rc;
    rc = check_arrays(power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, n_users, ld, d.P0);
    if (rc) return rc;
    if (n_beams < 1 || n_beams > 4096) return fail(DMK_ERR_INVALID_ARG, "n_beams=%d outside [1, 4096]", n_beams);
    if (d.has_time_axis) return fail(DMK_ERR_UNSUPPORTED, "beam amplitude maps take no time axis");
    if (d.rx_filter) return fail(DMK_ERR_UNSUPPORTED, "beam amplitude maps do not apply the receive low-pass filter");
    if (n_users == 0) return DMK_OK;
    if (d.K == 0) return fail(DMK_ERR_INVALID_ARG, "no selected subcarriers: the mean over subcarriers is undefined");
    if (!beams_c64 || !mean_abs) return fail(DMK_ERR_INVALID_ARG, "beams or output is NULL");
    if (n_users > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
    bind_arrays(d, power_dbw, phase_deg, delay_s, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, ue_rot_deg, nullptr, n_users, ld);
    d.fov_mask = fov_mask; d.valid_mask = valid_mask; d.clip_mask = clip_mask;
    if (d.K == 1 && d.subc_step == 0 && !d.subc) d.subc_step = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const DeviceState* dev = nullptr;
    rc = device_state(dev);
    if (rc) return rc;
    d.n_sms = dev->sms;
    // Production route: the packed-FP32 kernel in beam mode (virtual TX panel 1 x n_beams, dmk_fd.cuh: fd_fast_body<true>).
    // Needs an affine subcarrier selection and tables that fit; otherwise the generic tile version below runs.
    {
        const bool affine = d.subc_step != 0;
        DevDesc db = d;
        db.bs0 = 1; db.bs1 = n_beams; db.Mt = n_beams; db.M = d.Mr * n_beams;
        FastCfg cfg;
        BeamCfg bc;
        size_t fast_smem = 0;
        const int pc = d.P > 0 ? d.P : 1;
        auto take = [&](size_t bytes) { size_t o = fast_smem; fast_smem += (bytes + 15) & ~size_t(15); return (int)o; };
        cfg.pcap = pc;
        cfg.nA = (d.K + 15) / 16;
        cfg.off_W  = take((size_t)pc * kTKW * sizeof(float2));
        cfg.off_A  = take((size_t)8 * pc * 8 * sizeof(float4));
        cfg.off_tY = take((size_t)pc * sizeof(float2));
        cfg.off_tQ = take((size_t)pc * db.M * sizeof(float2));
        cfg.off_wA = take((size_t)pc * cfg.nA * sizeof(float2));
        cfg.off_wB = take((size_t)pc * 16 * sizeof(float2));
        bc.off_rows = take((size_t)db.M * sizeof(float));
        cfg.mul_mt  = db.Mt > 1 ? (unsigned)((0x100000000ULL + db.Mt - 1) / db.Mt) : 0u;
        cfg.mul_bs0 = 0u;
        bc.n_beams = n_beams; bc.bs0 = d.bs0; bc.bs1 = d.bs1;
        bc.F = reinterpret_cast<const float2*>(beams_c64);
        bc.out = mean_abs;
        const bool div_ok = (unsigned long long)db.M * (unsigned long long)db.Mt < 0xffffffffULL;
        // scratch tables [np][bs0|1], [np][bs1|1], [np][Mr] live in the W tile area, one staged codebook row [n_beams][bs0] in the
        // A strip area, and a thread accumulates at most 8 (beam, path) pairs
        const bool scratch_ok = (d.bs0 | 1) + (d.bs1 | 1) + d.Mr <= kTKW && (size_t)n_beams * d.bs0 * sizeof(float2) <= (size_t)8 * pc * 8 * sizeof(float4) &&
                                (long long)n_beams * pc <= 8LL * kFdThreads;
        if (affine && div_ok && fast_smem <= (size_t)kSmemFast && scratch_ok && desc->kernel_hint != DMK_KERNEL_TILE) {
            bf_fast_kernel<<<(unsigned)n_users, kFdThreads, fast_smem, st>>>(db, cfg, bc);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "bf_fast_kernel launch");
            g_launches.fetch_add(1);
            snprintf(g_kernel, sizeof(g_kernel), "bf_fast_kernel<64x256,ffma2> grid=%lld beams=%d smem=%zu", (long long)n_users, n_beams, fast_smem);
            return DMK_OK;
        }
    }
    BfCfg c;
    c.n_beams = n_beams;
    c.F = reinterpret_cast<const float2*>(beams_c64);
    c.out = mean_abs;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
    c.off_G    = take((size_t)n_beams * kMaxPaths * sizeof(float2));
    c.off_rows = take((size_t)d.Mr * n_beams * sizeof(float));
    c.off_tY   = take((size_t)kMaxPaths * d.bs0 * sizeof(float2));
    c.off_tZ   = take((size_t)kMaxPaths * d.bs1 * sizeof(float2));
    c.off_aR   = take((size_t)kMaxPaths * d.Mr * sizeof(float2));
    const size_t smem = (size_t)kMaxPaths * (kTK + kTM) * sizeof(float2) + off;
    if (smem > (size_t)kSmemTile) return fail(DMK_ERR_UNSUPPORTED, "beam tables need %zu bytes of shared memory (> 200 KB): fewer beams or a smaller panel", smem);
    bf_kernel<<<(unsigned)n_users, kFdThreads, smem, st>>>(d, c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "bf_kernel launch");
    g_launches.fetch_add(1);
    snprintf(g_kernel, sizeof(g_kernel), "bf_kernel<64x128> grid=%lld beams=%d smem=%zu", (long long)n_users, n_beams, smem);
    return DMK_OK;
}

int dmk_path_prologue(const dmk_desc* desc, const float* power_dbw, const float* aoa_az_deg, const float* aoa_el_deg,
                      const float* aod_az_deg, const float* aod_el_deg, const double* ue_rot_deg, int64_t n_users,
                      int32_t ld, double* angles_rot, double* power_gain, uint8_t* fov_mask, void* cuda_stream)
{
    using namespace dmk;
    DevDesc d;
    int rc = build_desc(desc, false, d);
    if (rc) return rc;
    rc = check_arrays(power_dbw, power_dbw, power_dbw, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, n_users, ld, d.P0);
    if (rc) return rc;
    if (n_users == 0) return DMK_OK;
    // phase/delay are not needed for the by-products: alias them to power (finite or NaN, never dereferenced wrongly)
    bind_arrays(d, power_dbw, power_dbw, power_dbw, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, ue_rot_deg, nullptr, n_users, ld);
    d.fov_mask = fov_mask;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const long long total = n_users * d.P0;
    const long long grid = (total + 255) / 256;
    if (grid > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
    prologue_kernel<<<(unsigned)grid, 256, 0, st>>>(d, angles_rot, power_gain);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "prologue_kernel launch");
    g_launches.fetch_add(1);
    return DMK_OK;
}

int dmk_user_byproducts(const dmk_desc* desc, const float* power_dbw, const float* phase_deg,
                        const float* aoa_az_deg, const float* aoa_el_deg, const float* aod_az_deg, const float* aod_el_deg,
                        const float* inter, const double* ue_rot_deg, int64_t n_users, int32_t ld,
                        int32_t* num_paths, int32_t* los, float* pathloss_coherent, float* pathloss_noncoherent, void* cuda_stream)
{
    using namespace dmk;
    DevDesc d;
    int rc = build_desc(desc, false, d);
    if (rc) return rc;
    rc = check_arrays(power_dbw, phase_deg, power_dbw, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, n_users, ld, d.P0);
    if (rc) return rc;
    if (n_users == 0) return DMK_OK;
    if (los && !inter) return fail(DMK_ERR_INVALID_ARG, "los needs the interaction codes");
    // delay is not needed for the by-products: alias it to power (never dereferenced by the time-domain gain chain)
    bind_arrays(d, power_dbw, phase_deg, power_dbw, aoa_az_deg, aoa_el_deg, aod_az_deg, aod_el_deg, ue_rot_deg, nullptr, n_users, ld);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const long long grid = (n_users * 32 + 255) / 256;
    if (grid > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "grid too large: split the user range");
    user_byproducts_kernel<<<(unsigned)grid, 256, 0, st>>>(d, inter, num_paths, los, pathloss_coherent, pathloss_noncoherent);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "user_byproducts_kernel launch");
    g_launches.fetch_add(1);
    return DMK_OK;
}

int dmk_np_sincosf(const float* x, float* s, float* c, int64_t n, void* cuda_stream)
{
    if (n < 0) return fail(DMK_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return DMK_OK;
    if (!x || !s || !c) return fail(DMK_ERR_INVALID_ARG, "NULL pointer");
    const long long grid = (n + 255) / 256;
    if (grid > 0x7fffffffLL) return fail(DMK_ERR_INVALID_ARG, "n too large");
    dmk::np_sincosf_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(x, s, c, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "np_sincosf_kernel launch");
    g_launches.fetch_add(1);
    return DMK_OK;
}

}  // extern "C"
