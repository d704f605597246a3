// dmk_fd_ws.cuh -- FD channel kernel on tcgen05, persistent and warp-specialised (the production tensor-core kernel).
//
// Same arithmetic as fd_tc_kernel (dmk_fd_tc.cuh: transposed tile, FP16 hi/lo operand split, FP32 accumulation in
// TMEM, per-user operand scale); what changes is who does what and when.  Phase traces of the one-CTA-per-user kernel
// on the city-scale shape (64 x 1024 x 8 B = 512 KB per user; tools/tc_trace.py, profiles/README.md) showed
//   * 20 % of a CTA's life in the float64 prologue + table building (no stores),
//   * every stage serialised as  build operands -> MMA -> drain, with all eight warps in the same phase, so the SM's
//     store stream stops whenever both resident CTAs build,
//   * the MMA batch itself taking 3-8 k cycles because a lone thread under `if (tid == 256)` makes the compiler wrap
//     every tcgen05.mma in an ELECT / R2UR.BROADCAST waterfall (~30 dependent instructions per MMA).
// Roles here (320 threads, 2 persistent CTAs per SM, users drawn from a device-side ticket counter):
//   warp 9      helper   ticket -> float64 prologue of the NEXT user (lanes = path columns) -> its phasor tables
//   warps 4-7   builders A/B operand tiles of stage g+1 (K-major SWIZZLE_128B, FP16 hi/lo) as soon as MMA(g) has read its operands
//   warp 8      issuer   elect.sync lane issues the tcgen05.mma batch of stage g, commits to mma_done[g & 1]
//   warps 0-3   drain    TMEM lane quarter q = warp: tcgen05.ld -> scale -> 128-byte row-segment stores of stage g-1
// All hand-offs are mbarriers (no CTA-wide barrier in the steady state); the accumulator is double-buffered in TMEM,
// the table set and the user record are double-buffered in shared memory.  Stores therefore never wait for operand
// building or for the prologue -- only for HBM.
#pragma once
#include "dmk_fd_tc.cuh"

namespace dmk {

constexpr int kWsConsumers = 288;  // warps 0-8: threads that read a user record (and arrive on ub_empty)
constexpr int kWsDrain0   = 0;     // warps 0-3
constexpr int kWsBuild0   = 4;     // warps 4-7
constexpr int kWsIssuer   = 8;
constexpr int kWsHelper0  = 9;     // warps 9 .. 9 + H - 1
constexpr int kWsBuilders = 128;
constexpr int kWsMaxHelpers = 4;
// user buffers per helper warp: two (it prepares user i + H while user i is consumed) or, where shared memory allows, three --
// users are consumed in ticket order, so a helper that is slow on a 25-path user holds up the consumers while the other helpers
// sit on finished users; a third buffer lets them run further ahead
constexpr int kWsMaxBufs = 12;

struct WsBars {
    uint64_t ub_full[kWsMaxBufs], ub_empty[kWsMaxBufs];   // helper -> everyone (1 arrival) ; everyone -> helper (288 arrivals)
    uint64_t op_full;                   // builders -> issuer (128 arrivals): operand tiles of the next stage are in smem
    uint64_t mma_done[2];               // tcgen05.commit: accumulator g & 1 complete, operand tiles free again
    uint64_t acc_empty[2];              // drain warps -> issuer (128 arrivals): accumulator g & 1 has been read out
};

__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
// Wait for the phase with the given parity to complete.  A lost arrival must fail loudly rather than hang the device, but a
// legitimately slow neighbour (time-sliced GPU, debugger) must not be killed: the watchdog runs on the global nanosecond
// timer and only fires after 30 s without progress on this barrier.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    const uint32_t bar = smem_u32(b);
    uint32_t ok = 0;
    unsigned long long t0 = 0;
    for (unsigned spins = 0; ; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if ((spins & 4095u) == 4095u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 30000000000ULL) __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

// One complex value per path slot x 4 -> FP16 hi/lo, written to the row pair (2c: Re H, 2c+1: Im H) of a B tile:
// row 2c holds (w.x, -w.y), row 2c+1 holds (w.y, w.x); the second row's halves are a sign flip / swap of the first's.
__device__ __forceinline__ void st_split8_f16_rowpair(unsigned char* hi, unsigned char* lo, int off_re, int off_im, const float2 (&w)[4])
{
    uint32_t hr[4], lr[4], hi_[4], li[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(w[i].x, w[i].y);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(w[i].x - back.x, w[i].y - back.y);
        const uint32_t h = *reinterpret_cast<const uint32_t*>(&hh), l = *reinterpret_cast<const uint32_t*>(&ll);
        hr[i] = h ^ 0x80000000u;  lr[i] = l ^ 0x80000000u;                    // (x, -y)
        hi_[i] = __byte_perm(h, 0, 0x1032); li[i] = __byte_perm(l, 0, 0x1032);  // (y, x)
    }
    *reinterpret_cast<uint4*>(hi + off_re) = make_uint4(hr[0], hr[1], hr[2], hr[3]);
    *reinterpret_cast<uint4*>(lo + off_re) = make_uint4(lr[0], lr[1], lr[2], lr[3]);
    *reinterpret_cast<uint4*>(hi + off_im) = make_uint4(hi_[0], hi_[1], hi_[2], hi_[3]);
    *reinterpret_cast<uint4*>(lo + off_im) = make_uint4(li[0], li[1], li[2], li[3]);
}

// Flat-chunk formulation (round 2).  A user's output [M rows][K columns] complex64 is a flat array of 512-byte CHUNKS: chunk
// c = m * S + seg holds the 64 subcarriers (128 floats) of segment seg of antenna row m, S = K / 64 (K not a multiple of 64:
// S = ceil(K / 64), the last chunk of a row is cut off -- ws_store_chunks_rag).  With the delay phasor split as
//     W[p, 64 seg + j] = wS[p, seg] * wF[p, j],        wS = exp(-j 2 pi wcyc (start + 64 step seg)),  wF = exp(-j 2 pi wcyc step j)
// the coarse factor moves to the antenna side:  H[c, j] = sum_p (A[m, p] wS[p, seg]) wF[p, j].  One tcgen05 stage then computes
//     D[2j + s, n] = sum_{p, e} Mside[2j + s, 2p + e] * Nside[n, 2p + e]          n = chunk c0 + n, up to 128 chunks per stage
// where Mside (the 2x2 real form of wF, 128 rows) is built ONCE PER USER and Nside per stage.  TMEM column n of the accumulator is
// chunk c0 + n and lane 2j + s its float: the accumulator read out column by column IS a contiguous 64 KB piece of the output, so
// every CTA writes one linear stream (7.4 TB/s in the pure-store micro-benchmark, tools/micro/store_rate3.cu, against 6.5-6.9 for
// the row-strided tile of round 1), and the per-column-stage B tile of round 1 is gone.
struct WsCfg {
    int off_N, off_M;                                   // operand tiles: byte offsets from the 1024-aligned base, hi then lo
    int off_tab, tab_bytes;                             // user buffers [TcUserBuf][tables]
    int off_tY, off_tQ, off_wB, off_wL, off_wS;         // byte offsets inside a table buffer
    int sY, sQ, sB, sL, sS;                             // per-path table strides (float2 units), odd
    int S, n_chunks, n_stages;                          // segments per row, chunks per user, ceil(n_chunks / 128)
    int bufs_per_helper;                                // 2 or 3 user buffers per helper warp
    int rag, row_floats, last_valid;                    // K % 64 != 0: S = ceil(K / 64); floats per antenna row (2 K); floats of a row's last chunk
    unsigned mul_mt, mul_bs0, mul_s;                    // ceil(2^32 / d) reciprocals (0: d == 1)
};

// Helper warp: ticket -> prologue -> per-user tables of one buffer.  lanes = path columns, then lanes = table entries.
// Row np of every table is zero-filled: the operand builders read it (index min(p, np)) for the padding slots.
// kSfu: table phasors on the SFU (phasor_cycles_sfu, <= 3.6e-7 absolute per unit phasor -- the class of the FP16 hi/lo operand
// split the entries go through, 2^-22) instead of the float32 polynomial.  Used by the four-helper instantiation only, i.e. for
// per-user outputs below ~400 KB, where this warp -- the per-user serial path -- bounds the kernel.
template <bool kSfu>
__device__ __forceinline__ float2 ws_phasor(double cyc) { return kSfu ? phasor_cycles_sfu(cyc) : phasor_cycles(cyc); }

template <bool kSfu>
__device__ __noinline__ void ws_helper_prepare(const DevDesc& d, const WsCfg& cfg, int ksplit, unsigned int n_items, unsigned int n_draw_last,
                                               unsigned int* ticket, TcUserBuf& ub, unsigned char* tab, int lane)
{
    unsigned int t = 0;
    if (lane == 0) {
        t = atomicAdd(ticket, 1u);
        if (t == n_draw_last) atomicExch(ticket, 0u);      // every helper warp of every CTA draws exactly one ticket >= n_items: this is the last draw
        ub.item = t;
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= n_items) return;
    const long long user = t / (unsigned)ksplit;
    FdShared& sh = ub.sh;

    PathState st;
    const bool active = lane < d.P0;
    st.contrib = false; st.valid = false; st.fov = true; st.over = false;
    st.c = make_float2(0.f, 0.f);
    if (active) {
        SideOut s0, s1;
        GainOut g;
        if (prologue_needs_angles(d)) { prologue_side<true>(d, user, lane, 0, s0, d.Mt > 1);  prologue_side<true>(d, user, lane, 1, s1, d.Mr > 1); }
        else                          { prologue_side<false>(d, user, lane, 0, s0, d.Mt > 1); prologue_side<false>(d, user, lane, 1, s1, d.Mr > 1); }
        prologue_gain<true>(d, user, lane, g);
        prologue_combine<true>(d, s0, s1, g, st);
    }
    const bool contrib = active && st.contrib;
    const unsigned ballot = __ballot_sync(0xffffffffu, contrib);
    const int np = __popc(ballot);
    if (contrib) {
        const int j = __popc(ballot & ((1u << lane) - 1u));
        sh.c[j] = st.c; sh.wcyc[j] = st.wcyc; sh.fd[j] = st.fd;
        sh.u[0][j] = st.u[0]; sh.v[0][j] = st.v[0];
        sh.u[1][j] = st.u[1]; sh.v[1][j] = st.v[1];
    }
    if (lane == 0) sh.np = np;
    // The masks are outputs: they leave the SM with the drain warps, the only role that orders itself after the previous launch
    // (griddepcontrol.wait); everything the helper, the builders and the issuer do touches shared / tensor memory and inputs only.
    ub.m_fov[lane]   = st.fov ? 1 : 0;
    ub.m_valid[lane] = st.valid ? 1 : 0;
    ub.m_clip[lane]  = (st.valid && st.over) ? 1 : 0;
    if (np == 0) return;
    // per-user operand scale: largest |c_p| component (FP16 operands live in [-1, 1])
    float mx = contrib ? fmaxf(fabsf(st.c.x), fabsf(st.c.y)) : 0.f;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) ub.scale = mx;
    const float inv_scale = 1.0f / mx;
    __syncwarp();

    float2* tY = reinterpret_cast<float2*>(tab + cfg.off_tY);
    float2* tQ = reinterpret_cast<float2*>(tab + cfg.off_tQ);
    float2* wB = reinterpret_cast<float2*>(tab + cfg.off_wB);
    float2* wL = reinterpret_cast<float2*>(tab + cfg.off_wL);
    float2* wS = reinterpret_cast<float2*>(tab + cfg.off_wS);
    const int bs0 = d.bs0, bs1 = d.bs1, nq = d.Mr * d.bs1, S = cfg.S;
    for (int e = lane; e < np * bs0; e += 32) {
        const int p = e / bs0, y = e - p * bs0;
        tY[p * cfg.sY + y] = ws_phasor<kSfu>((double)y * sh.u[0][p]);
    }
    for (int e = lane; e < np * nq; e += 32) {
        const int p = e / nq, q = e - p * nq;
        const int r = q / bs1, z = q - r * bs1;
        const int yr = r % d.ue0, zr = r / d.ue0;
        const float2 cs = make_float2(sh.c[p].x * inv_scale, sh.c[p].y * inv_scale);
        tQ[p * cfg.sQ + q] = cmul(cs, ws_phasor<kSfu>((double)z * sh.v[0][p] + (double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
    }
    for (int e = lane; e < np * 20; e += 32) {                          // fine delay phasors: wB[b], b < 16, and wL[a], a < 4 (j = 16 a + b)
        const int p = e / 20, b = e - p * 20;
        if (b < 16) wB[p * cfg.sB + b] = ws_phasor<kSfu>(-(sh.wcyc[p] * (double)(d.subc_step * b)));
        else        wL[p * cfg.sL + (b - 16)] = ws_phasor<kSfu>(-(sh.wcyc[p] * (double)(d.subc_step * 16 * (b - 16))));
    }
    for (int e = lane; e < np * S; e += 32) {                           // coarse delay phasors, one per 64-subcarrier segment
        const int p = e / S, sg = e - p * S;
        wS[p * cfg.sS + sg] = ws_phasor<kSfu>(-(sh.wcyc[p] * ((double)d.subc_start + (double)d.subc_step * 64.0 * (double)sg)));
    }
    // row np of every table is the zero row: the operand builders read it for the padding slots np .. nslot-1
    const float2 zero = make_float2(0.f, 0.f);
    for (int e = lane; e < bs0; e += 32) tY[np * cfg.sY + e] = zero;
    for (int e = lane; e < nq; e += 32)  tQ[np * cfg.sQ + e] = zero;
    for (int e = lane; e < 16; e += 32)  wB[np * cfg.sB + e] = zero;
    for (int e = lane; e < 4; e += 32)   wL[np * cfg.sL + e] = zero;
    for (int e = lane; e < S; e += 32)   wS[np * cfg.sS + e] = zero;
}

// Drain 32 accumulator columns = 32 consecutive chunks of the output: register i holds, across the lanes of the warp, floats
// q*32 .. q*32+31 of chunk (first + i).  Chunks are 512 bytes apart, so every store is base + i * 512: an immediate offset, no
// address arithmetic between the stores (the row-pitch version of round 1 spent two instructions per store on it).
__device__ __forceinline__ void ws_store_chunks(uint32_t taddr, float* out, int first, int rows, float scale)
{
    uint32_t v[32];
    tmem_ld<32>(taddr, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (first + 32 <= rows) {
        #pragma unroll
        for (int i = 0; i < 32; ++i) __stcs(out + i * kTcN, __uint_as_float(v[i]) * scale);    // * scale: undo the per-user operand scale
    } else {
        #pragma unroll
        for (int i = 0; i < 32; ++i)
            if (first + i < rows) __stcs(out + i * kTcN, __uint_as_float(v[i]) * scale);
    }
}

// K not a multiple of 64 (12 x n resource blocks: 300, 600, 624, 1200 ... subcarriers): a row has S = ceil(K / 64) chunks, the last one is
// cut off at column K, and the chunks of consecutive rows are no longer 512 bytes apart (row pitch 2 K floats).  (m, seg) of the 32
// chunks run as warp-uniform counters; `out` points at float q * 32 + lane of chunk `c0`, `lane_ok` = this float exists in a last chunk.
__device__ __forceinline__ void ws_store_chunks_rag(const uint32_t (&v)[32], float* out, int seg, int S, int delta, int left, bool lane_ok, float scale)
{
    #pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < left && (seg < S - 1 || lane_ok)) __stcs(out, __uint_as_float(v[i]) * scale);
        out += kTcN;
        if (++seg == S) { seg = 0; out += delta; }           // next antenna row: + 2 K - 128 S floats
    }
}

// Consumers (drain, builders, issuer) walk the users in the order it = 0, 1, 2, ...: user `it` is prepared by helper it % H into
// buffer it % (2H) (every helper owns two buffers).  A helper that draws a ticket >= n_items publishes it as a sentinel and stops;
// its slots are skipped from then on, and the walk ends when every helper of the CTA has stopped.
template <int H>
__device__ __forceinline__ bool ws_next_user(WsBars& bars, const unsigned char* bufs, int buf_stride, unsigned n_items,
                                             unsigned& it, unsigned& done, int& b, unsigned R)
{
    constexpr unsigned kAll = (1u << H) - 1u;
    for (;;) {
        if (done == kAll) return false;
        const unsigned h = it % H;
        if (!((done >> h) & 1u)) {
            b = (int)(it % R);
            mbar_wait(&bars.ub_full[b], (it / R) & 1u);
            if (reinterpret_cast<const TcUserBuf*>(bufs + (size_t)b * buf_stride)->item < n_items) return true;
            done |= 1u << h;
        }
        ++it;
    }
}

template <int H>        // helper warps per CTA: 1 (two CTAs per SM, large per-user outputs) or 4 (one CTA per SM, helper-bound shapes)
__global__ void __launch_bounds__((9 + H) * 32, H <= 2 ? 2 : 1)
fd_ws_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ WsCfg cfg, const int ksplit,
             const unsigned int n_items, unsigned int* ticket, const int pdl_wait)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ WsBars bars;
    // Programmatic dependent launch.  The next launch may take SMs as our CTAs retire; unless the caller declared this launch
    // independent, the drain warps -- the only role that writes global memory -- order themselves after the previous grid before
    // their first store, while prologue, tables, operand tiles and the first MMAs of this launch already run under its tail.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* sNhi = sm + cfg.off_N;                    // [128 chunk rows][128 B]: 32 path slots x (re, im) fp16, K-major SWIZZLE_128B
    unsigned char* sNlo = sNhi + kTcN * 128;
    unsigned char* sMhi = sm + cfg.off_M;                    // [128 rows 2j + s][128 B]: fine delay phasors of the user, 2x2 real form
    unsigned char* sMlo = sMhi + kTcN * 128;
    unsigned char* bufs = sm + cfg.off_tab;                  // bufs_per_helper * H buffers of cfg.tab_bytes: [TcUserBuf][tables]
    constexpr int kUb = (int)((sizeof(TcUserBuf) + 15) & ~size_t(15));
    auto user_buf = [&](int b) -> TcUserBuf& { return *reinterpret_cast<TcUserBuf*>(bufs + (size_t)b * cfg.tab_bytes); };
    auto user_tab = [&](int b) -> unsigned char* { return bufs + (size_t)b * cfg.tab_bytes + kUb; };

    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(2 * 128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 64) {
        for (int b = 0; b < cfg.bufs_per_helper * H; ++b) { mbar_init(&bars.ub_full[b], 1); mbar_init(&bars.ub_empty[b], kWsConsumers); }
        mbar_init(&bars.op_full, kWsBuilders);
        mbar_init(&bars.mma_done[0], 1);  mbar_init(&bars.mma_done[1], 1);
        mbar_init(&bars.acc_empty[0], 128); mbar_init(&bars.acc_empty[1], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int n_chunks = cfg.n_chunks, n_stages = cfg.n_stages;

    if (warp >= kWsHelper0) {
        // ------------------------------------------------------------------------------------------ helpers
        const int h = warp - kWsHelper0;
        const unsigned int n_draw_last = n_items + gridDim.x * (unsigned)H - 1u;
        for (unsigned k = 0;; ++k) {
            const unsigned B = (unsigned)cfg.bufs_per_helper;
            const int b = h + (int)(k % B) * H;                           // this helper's buffers, round robin
            mbar_wait(&bars.ub_empty[b], ((k / B) & 1u) ^ 1u);            // the readers of this buffer's previous user are done
            ws_helper_prepare<false>(d, cfg, ksplit, n_items, n_draw_last, ticket, user_buf(b), user_tab(b), lane);
            __syncwarp();
            const unsigned int item = user_buf(b).item;
            if (lane == 0) mbar_arrive(&bars.ub_full[b]);
            if (item >= n_items) break;
        }
    } else if (warp >= kWsBuild0 && warp < kWsIssuer) {
        // ------------------------------------------------------------------------------------------ operand builders
        const int bt = tid - kWsBuild0 * 32;                  // 0..127: N-side row (chunk of the stage) / M-side subcarrier + group
        const int n_off0 = (bt >> 3) * 1024 + (bt & 7) * 128;
        const int n_sw = bt & 7;
        const int b_col  = bt & 63;
        const int b_grp  = bt >> 6;                           // 0..1
        const int b_row0 = 2 * b_col;                         // rows 2j (Re H) and 2j + 1 (Im H)
        const int b_off0 = (b_row0 >> 3) * 1024 + (b_row0 & 7) * 128;
        const int b_sw0 = b_row0 & 7, b_sw1 = (b_row0 + 1) & 7;
        unsigned g = 0;                                       // global stage counter (identical in every role)
        unsigned it = 0, done = 0;
        int cur = 0;
        for (; ws_next_user<H>(bars, bufs, cfg.tab_bytes, n_items, it, done, cur, (unsigned)(cfg.bufs_per_helper * H)); ++it) {
            const TcUserBuf& ub = user_buf(cur);
            const unsigned int item = ub.item;
            const int ks = (int)(item % (unsigned)ksplit);
            const int np = ub.sh.np;
            if (np > 0) {
                const unsigned char* tab = user_tab(cur);
                const float2* tY = reinterpret_cast<const float2*>(tab + cfg.off_tY);
                const float2* tQ = reinterpret_cast<const float2*>(tab + cfg.off_tQ);
                const float2* wB = reinterpret_cast<const float2*>(tab + cfg.off_wB);
                const float2* wL = reinterpret_cast<const float2*>(tab + cfg.off_wL);
                const float2* wS = reinterpret_cast<const float2*>(tab + cfg.off_wS);
                const int nslot = ((np + 7) >> 3) << 3;       // slots the tensor core reads (table row np is the zero row)
                const int nq4 = nslot >> 2;
                bool m_valid = false;
                for (int stg = ks; stg < n_stages; stg += ksplit, ++g) {
                    if (g > 0) mbar_wait(&bars.mma_done[(g - 1) & 1], ((g - 1) >> 1) & 1u);     // MMA(g-1) has read the operand tiles
                    if (!m_valid) {
                        // ---- M side, once per user: rows 2j -> (Re wF, -Im wF), 2j + 1 -> (Im wF, Re wF), wF[p, j] = wL[p, j >> 4] wB[p, j & 15]
                        const float2* wl0 = wL + (b_col >> 4);
                        const float2* wb0 = wB + (b_col & 15);
                        #pragma unroll 1
                        for (int qd = b_grp; qd < nq4; qd += 2) {
                            float2 l4[4], b4[4], w[4];
                            #pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int p = min(qd * 4 + i, np);
                                l4[i] = wl0[p * cfg.sL]; b4[i] = wb0[p * cfg.sB];
                            }
                            #pragma unroll
                            for (int i = 0; i < 4; ++i) w[i] = cmul(l4[i], b4[i]);
                            st_split8_f16_rowpair(sMhi, sMlo, b_off0 + (((qd ^ b_sw0) & 7) << 4), b_off0 + 128 + (((qd ^ b_sw1) & 7) << 4), w);
                        }
                        m_valid = true;
                    }
                    // ---- N side: row n = chunk c0 + n = (antenna row m, segment seg): A[m, p] * wS[p, seg]
                    {
                        const unsigned c = (unsigned)(stg * kTcN + bt);
                        const bool ok = c < (unsigned)n_chunks;
                        int a_q = 0, a_y = 0, sg = 0;
                        if (ok) {
                            const unsigned mm = cfg.mul_s ? __umulhi(c, cfg.mul_s) : c;              // antenna row
                            sg = (int)(c - mm * (unsigned)cfg.S);
                            const unsigned rr = cfg.mul_mt ? __umulhi(mm, cfg.mul_mt) : mm;
                            const unsigned t = mm - rr * (unsigned)d.Mt;
                            const unsigned zt = cfg.mul_bs0 ? __umulhi(t, cfg.mul_bs0) : t;
                            a_y = (int)(t - zt * (unsigned)d.bs0);
                            a_q = (int)(rr * (unsigned)d.bs1 + zt);
                        }
                        const float2* q0 = tQ + a_q;
                        const float2* y0 = tY + a_y;
                        const float2* s0 = wS + sg;
                        // rows past the user's last chunk read the zero row (index np) of every table: no branch in the loop, the
                        // twelve table loads of a slot quad are independent and issue back to back
                        const int p_cap = ok ? np : 0, p_add = ok ? 0 : np;
                        #pragma unroll 2
                        for (int qd = 0; qd < nq4; ++qd) {
                            float2 tq[4], ty[4], ts[4];
                            #pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int p = min(qd * 4 + i, p_cap) + p_add;
                                tq[i] = q0[p * cfg.sQ]; ty[i] = y0[p * cfg.sY]; ts[i] = s0[p * cfg.sS];
                            }
                            float2 a[4];
                            #pragma unroll
                            for (int i = 0; i < 4; ++i) a[i] = cmul(cmul(tq[i], ty[i]), ts[i]);
                            st_split8_f16(sNhi, sNlo, n_off0 + (((qd ^ n_sw) & 7) << 4), a);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&bars.op_full);
                }
            }
            mbar_arrive(&bars.ub_empty[cur]);
        }
    } else if (warp == kWsIssuer) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        const uint64_t dNhi = umma_desc_kmajor_sw128(smem_u32(sNhi)), dNlo = umma_desc_kmajor_sw128(smem_u32(sNlo));
        const uint64_t dMhi = umma_desc_kmajor_sw128(smem_u32(sMhi)), dMlo = umma_desc_kmajor_sw128(smem_u32(sMlo));
        unsigned g = 0;
        unsigned it = 0, done = 0;
        int cur = 0;
        for (; ws_next_user<H>(bars, bufs, cfg.tab_bytes, n_items, it, done, cur, (unsigned)(cfg.bufs_per_helper * H)); ++it) {
            const TcUserBuf& ub = user_buf(cur);
            const unsigned int item = ub.item;
            const int ks = (int)(item % (unsigned)ksplit);
            const int np = ub.sh.np;
            if (np > 0) {
                const int ksteps = (np + 7) >> 3;                     // 16 fp16 (8 path slots) per MMA
                for (int stg = ks; stg < n_stages; stg += ksplit, ++g) {
                    const unsigned ab = g & 1;
                    // tcgen05 N = chunks of this stage, rounded up to the instruction granularity (16 for M = 128)
                    const int n_here = min(kTcN, (n_chunks - stg * kTcN + 15) & ~15);
                    // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = chunks, M = 128 floats of a chunk
                    const uint32_t idesc = (1u << 4) | ((uint32_t)(n_here >> 3) << 17) | ((uint32_t)(kTcN >> 4) << 24);
                    mbar_wait(&bars.op_full, g & 1u);                                 // operand tiles of stage g are in smem
                    mbar_wait(&bars.acc_empty[ab], ((g >> 1) & 1u) ^ 1u);              // accumulator ab drained (stage g-2)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t acc = tmem_base + ab * (uint32_t)kTcN;
                        #pragma unroll
                        for (int s = 0; s < 3; ++s) {                                           // hi*hi, lo*hi, hi*lo
                            const uint64_t dn = (s == 2) ? dNlo : dNhi;
                            const uint64_t dm = (s == 1) ? dMlo : dMhi;
                            #pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (kk < ksteps) {
                                    const uint32_t accum = (s | kk) != 0;
                                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                                 :: "r"(acc), "l"(dm + 2 * kk), "l"(dn + 2 * kk), "r"(idesc), "r"(accum) : "memory");
                                }
                            }
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     :: "r"(smem_u32(&bars.mma_done[ab])) : "memory");
                    }
                    __syncwarp();
                }
            }
            mbar_arrive(&bars.ub_empty[cur]);
        }
    } else {
        // ------------------------------------------------------------------------------------------ drain warps 0-3
        const int q = warp;                                   // TMEM lane quarter = 32 floats (128 bytes) of every chunk
#ifdef DMK_TC_TRACE
        unsigned long long tr_t0 = 0, tr_first = 0, tr_users = 0;
        if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_t0));
#endif
        if (pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");     // everything before this launch has completed and is visible
        unsigned g = 0;
        unsigned it = 0, done = 0;
        int cur = 0;
        for (; ws_next_user<H>(bars, bufs, cfg.tab_bytes, n_items, it, done, cur, (unsigned)(cfg.bufs_per_helper * H)); ++it) {
            const TcUserBuf& ub = user_buf(cur);
            const unsigned int item = ub.item;
            const long long user = item / (unsigned)ksplit;
            const int ks = (int)(item % (unsigned)ksplit);
            const int np = ub.sh.np;
            const float scale = ub.scale;
            float* out_u = reinterpret_cast<float*>(d.out) + user * (cfg.rag ? (long long)d.M * cfg.row_floats : (long long)n_chunks * kTcN);
#ifdef DMK_TC_TRACE
            if (tid == 0) { if (!tr_first) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_first)); ++tr_users; }
#endif
            if (q == 0 && ks == 0 && lane < d.P0) {
                const long long o = user * (long long)d.P0 + lane;
                if (d.fov_mask)   d.fov_mask[o]   = ub.m_fov[lane];
                if (d.valid_mask) d.valid_mask[o] = ub.m_valid[lane];
                if (d.clip_mask)  d.clip_mask[o]  = ub.m_clip[lane];
            }
            if (cfg.rag) {
                // cut-off last chunks: per-chunk addresses from running (row, segment) counters; zero users take the same route
                const bool lane_ok = q * 32 + lane < cfg.last_valid;
                const int delta = cfg.row_floats - cfg.S * kTcN;
                for (int stg = ks; stg < n_stages; stg += ksplit) {
                    const int rows = min(kTcN, n_chunks - stg * kTcN);
                    uint32_t taddr = 0; unsigned ab = 0;
                    if (np != 0) {
                        ab = g & 1;
                        mbar_wait(&bars.mma_done[ab], (g >> 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * (uint32_t)kTcN;
                    }
                    for (int r = 0; r < rows; r += 32) {
                        const unsigned c0 = (unsigned)(stg * kTcN + r);
                        const unsigned m0 = cfg.mul_s ? __umulhi(c0, cfg.mul_s) : c0;
                        const int seg0 = (int)(c0 - m0 * (unsigned)cfg.S);
                        uint32_t v[32];
                        if (np != 0) { tmem_ld<32>(taddr + r, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
                        else {
                            #pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0u;
                        }
                        ws_store_chunks_rag(v, out_u + (long long)m0 * cfg.row_floats + seg0 * kTcN + q * 32 + lane, seg0, cfg.S, delta, rows - r, lane_ok, np != 0 ? scale : 1.f);
                    }
                    if (np != 0) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(&bars.acc_empty[ab]);
                        ++g;
                    }
                }
            } else if (np == 0) {
                // users without contributing paths: zeros (channel.py:257,:269-271); a stage is 64 KB of contiguous output
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int stg = ks; stg < n_stages; stg += ksplit) {
                    const int n4 = min(kTcN, n_chunks - stg * kTcN) * (kTcN / 4);             // float4 of this stage
                    float4* o = reinterpret_cast<float4*>(out_u + (long long)stg * kTcN * kTcN);
                    for (int e = warp * 32 + lane; e < n4; e += 128) __stcs(o + e, z);
                }
            } else {
                for (int stg = ks; stg < n_stages; stg += ksplit, ++g) {
                    const unsigned ab = g & 1;
                    const int rows = min(kTcN, n_chunks - stg * kTcN);                        // chunks of this stage
                    mbar_wait(&bars.mma_done[ab], (g >> 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * (uint32_t)kTcN;
                    float* o = out_u + (long long)stg * kTcN * kTcN + q * 32 + lane;        // chunk r of the stage: + r * 128 floats
                    for (int r = 0; r < rows; r += 32)
                        ws_store_chunks(taddr + r, o + r * kTcN, r, rows, scale);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&bars.acc_empty[ab]);
                }
            }
            mbar_arrive(&bars.ub_empty[cur]);
        }
#ifdef DMK_TC_TRACE
        if (tid == 0 && blockIdx.x < 296) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            const int slot = (int)((ticket - g_tc_ticket) % 3) * 1184 + blockIdx.x * 4;       // three consecutive launches keep their records
            g_tc_trace[slot + 0] = (long long)tr_t0; g_tc_trace[slot + 1] = (long long)tr_first;
            g_tc_trace[slot + 2] = (long long)t_end; g_tc_trace[slot + 3] = (long long)tr_users;
        }
#endif
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(2 * 128));
}

}  // namespace dmk
