// Fused front end of the batched PicketFence pipeline: ONE persistent CTA per frame does
//   frame statistics (min / max / sum / row + column sums / corner boxes / exact p0.5, p99.5 and median)
//   -> noise + inversion decision -> clamped row / column sums (orientation) -> leaf profile -> picket search.
// It replaces k_frame_stats + k_pf_decide + k_pf_clamp_sums + k_pf_profile of pf.cu on the fast path.
//
// Reference semantics: PFDicomImage._has_noise / check_inversion (picketfence.py:221-238, core/image.py:868-897),
// ground + normalize (picketfence.py:322-323), PicketFence.orientation (picketfence.py:1501-1526), picket search
// (picketfence.py:747-767).
//
// Why not a histogram: per-pixel shared-memory atomics are capped by the LSU at ~2 cycles per lane (~4 % of the HBM
// roofline) and collapse when many pixels share one value (clipped noise floors).  The exact order statistics are
// obtained Floyd-Rivest style instead:
//   T0   a 2048-pixel sample per frame, 16-step value bisection in registers -> WIDE bands around each target rank;
//   P    the pilot rows (every 16th row, ~6 % of the frame) are streamed with the wide bands: pixels below a band are
//        counted in registers, pixels inside go to a small shared-memory histogram (rare);
//   R    re-pivot: the pilot histogram gives NARROW bands (+-5 sigma of the rank estimate, ~1-2 % of the pixels);
//   M    the remaining rows are streamed with the narrow bands;
//   X    resolve: rank - (#pixels below the band) indexes the band histogram -> the exact order statistic.  If a target
//        rank falls outside its band (probability ~1e-6 per frame) the frame is flagged and re-run by the exact
//        histogram pipeline of pf.cu -- results are always exact.
// A band whose answer is provably the frame minimum / maximum (e.g. a clipped floor holding more pixels than the
// target rank) is switched off after the pilot; the proof (pilot count + global min / max) is checked at resolve time.
//
// Streaming structure: warp per row, each lane owns 4 x 8-pixel vectors (128-bit ld.global.nc.L1::no_allocate) of a
// 1024-pixel column strip, next row prefetched while the current one is processed.  Column sums live in registers
// (IDP.2A on the packed u16x2 words), row sums are one REDUX per row, min / max are VIMNMX3.U16x2, band tests use
// per-vector packed min / max to skip the low / high bands for almost every vector.
#include "pf_common.cuh"

namespace epid {

constexpr int FR_THREADS = 512;
constexpr int FR_WARPS = FR_THREADS / 32;
constexpr int FR_VPL = 2;                       // vectors per lane per strip
constexpr int FR_STRIP_VEC = 32 * FR_VPL;       // 128 vectors = 1024 pixels
constexpr int FR_CAP = 4096;                    // histogram bins per band
constexpr int FR_NB = 3;                        // bands: 0 = low pair, 1 = median pair, 2 = high pair (flipped domain)
constexpr int FR_T0 = 4;                        // sample pixels per thread
constexpr int FR_PILOT_STEP = 16, FR_PILOT_OFF = 8;
constexpr int FR_STAGE_WARPS = 8;

struct BandRT {          // register copy of one band: accept a + e <= v' <= a + width, v' = flip ? 65535 - v : v
    uint32_t a, acc_n, off, e;   // acc_n = number of accepted values, off = a - a0 (histogram index of v' == a)
};

struct FrontSh {
    uint32_t bis_lo[8], bis_hi[8], bis_cnt[8], bis_rank[8];
    uint32_t a0[FR_NB], b0[FR_NB];                       // pilot bands (histogram origins)
    uint32_t a1[FR_NB], b1[FR_NB], e1[FR_NB], on1[FR_NB], cand[FR_NB];
    uint32_t below_p[FR_NB], below_m[FR_NB];
    uint32_t mn, mx, pmn, pmx;
    unsigned long long corner;
    uint32_t red[40];
    uint32_t scan_total;
    int fallback;
    uint32_t found[8];
    FrameStats st;
};

__device__ __forceinline__ uint32_t dp2a(uint32_t w, uint32_t sel, uint32_t c) { return __dp2a_lo(w, sel, c); }

struct Acc {
    uint32_t cs[FR_VPL * 8];
    uint32_t mn2, mx2;
    uint32_t below[FR_NB], le[FR_NB];
};

__device__ __forceinline__ void band_px(uint32_t v, const BandRT& b, uint32_t* __restrict__ hist, uint32_t& below, uint32_t& le, bool e) {
    const int u = (int)v - (int)b.a;
    below += (uint32_t)u >> 31;
    if (e) le += (uint32_t)(u - 1) >> 31;
    if ((uint32_t)(u - (int)(e ? 1 : 0)) < b.acc_n) atomicAdd(&hist[u + b.off], 1u);
}

// exclusive block scan of one value per thread (FR_THREADS threads); returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* s_red, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_red[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = lane < FR_WARPS ? s_red[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < FR_WARPS) s_red[lane] = winc - w;
        if (lane == 31) s_red[32] = winc;
    }
    __syncthreads();
    const uint32_t excl = s_red[wid] + inc - v;
    *total = s_red[32];
    return excl;
}

// smallest histogram bin index x in [0, nb) with base + sum(hist[0..x]) > rank; nb if none.  All threads call.
__device__ inline uint32_t hist_find(const uint32_t* __restrict__ hist, uint32_t nb, uint32_t base, uint32_t rank, FrontSh* sh, int slot) {
    const int tid = threadIdx.x;
    const uint32_t per = (nb + FR_THREADS - 1) / FR_THREADS;
    const uint32_t lo = min(nb, tid * per), hi = min(nb, lo + per);
    uint32_t c = 0;
    for (uint32_t i = lo; i < hi; i++) c += hist[i];
    uint32_t total;
    const uint32_t excl = block_scan_excl(c, sh->red, &total);
    if (tid == 0) sh->found[slot] = nb;
    __syncthreads();
    if (rank >= base) {
        const uint32_t k = rank - base;
        if (k >= excl && k < excl + c) {
            uint32_t acc = excl;
            for (uint32_t i = lo; i < hi; i++) {
                const uint32_t h = hist[i];
                if (k < acc + h) { sh->found[slot] = i; break; }
                acc += h;
            }
        }
    } else if (tid == 0) {
        sh->found[slot] = 0xffffffffu;   // the rank lies below the band
    }
    __syncthreads();
    const uint32_t r = sh->found[slot];
    __syncthreads();
    return r;
}

__device__ inline uint32_t hist_sum_below(const uint32_t* __restrict__ hist, uint32_t nb, FrontSh* sh) {
    const int tid = threadIdx.x;
    uint32_t c = 0;
    for (uint32_t i = tid; i < nb; i += FR_THREADS) c += hist[i];
    c = warp_sum(c);
    __syncthreads();
    if (tid == 0) sh->scan_total = 0;
    __syncthreads();
    if ((tid & 31) == 0 && c) atomicAdd(&sh->scan_total, c);
    __syncthreads();
    const uint32_t r = sh->scan_total;
    __syncthreads();
    return r;
}

struct View {
    const uint16_t* __restrict__ f;   // view origin
    int pitch, H, W, mis, nvec;       // nvec = vectors of the aligned grid covering the view
};

__device__ __forceinline__ int pilot_rows(int H) { return H > FR_PILOT_OFF ? (H - FR_PILOT_OFF + FR_PILOT_STEP - 1) / FR_PILOT_STEP : 0; }
__device__ __forceinline__ int phase_row(int phase, int i) {
    return phase == 0 ? FR_PILOT_OFF + FR_PILOT_STEP * i : i + (i + (FR_PILOT_STEP - 1 - FR_PILOT_OFF)) / (FR_PILOT_STEP - 1);
}

// 8 pixels of one vector through one band (rare path: kept out of line so the hot loop stays small)
__device__ __noinline__ uint2 band_vec(uint4 q, uint32_t a, uint32_t acc_n, uint32_t off, uint32_t e, uint32_t flip, uint32_t* __restrict__ hist) {
    uint32_t below = 0, le = 0;
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    BandRT b;
    b.a = a; b.acc_n = acc_n; b.off = off; b.e = e;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const uint32_t lo = w[t] & 0xffffu, hi = w[t] >> 16;
        band_px(flip ? 65535u - lo : lo, b, hist, below, le, e != 0);
        band_px(flip ? 65535u - hi : hi, b, hist, below, le, e != 0);
    }
    return make_uint2(below, le);
}

// One streaming phase over one column strip.  phase 0: pilot rows, 1: the other rows.  Only vectors that lie completely
// inside the view are handled here; the (at most 14) edge columns of a misaligned view go through edge_columns().
template <bool MED_E>
__device__ __forceinline__ void stream_phase(const View& vw, int strip, int phase, const BandRT (&bd)[FR_NB], bool lo_on, bool hi_on,
                                             uint32_t* __restrict__ hist, uint32_t* __restrict__ rowsum_sm, Acc& acc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int np = pilot_rows(vw.H);
    const int nrows = phase == 0 ? np : vw.H - np;
    bool full[FR_VPL];
#pragma unroll
    for (int k = 0; k < FR_VPL; k++) {
        const int j = strip * FR_STRIP_VEC + lane + 32 * k;
        const int c0 = j * 8 - vw.mis;
        full[k] = j < vw.nvec && c0 >= 0 && c0 + 8 <= vw.W;
    }
    // packed skip-test constants
    const uint32_t lo_b1 = bd[0].a + bd[0].acc_n + bd[0].e;                 // b + 1
    const bool lo_always = lo_b1 > 65535u;
    const uint32_t lo_c = min(lo_b1, 65535u) * 0x00010001u;
    // flipped band: v' = 65535 - v <= b'  <=>  v >= 65535 - b'
    const uint32_t hb = bd[2].a + bd[2].acc_n + bd[2].e - 1u;               // b'
    const bool hi_always = hb >= 65535u;
    const uint32_t hi_c = (hi_always ? 0u : 65535u - hb - 1u) * 0x00010001u;   // (lim - 1) in both halves

    const uint16_t* __restrict__ base = vw.f - vw.mis + (size_t)strip * FR_STRIP_VEC * 8;
    uint4 q[FR_VPL], qn[FR_VPL];
    int ri = wid;
    if (ri < nrows) {
        const uint16_t* rp = base + (size_t)phase_row(phase, ri) * vw.pitch;
#pragma unroll
        for (int k = 0; k < FR_VPL; k++) q[k] = full[k] ? ldg_stream16(rp + (lane + 32 * k) * 8) : make_uint4(0, 0, 0, 0);
    }
    while (ri < nrows) {
        const int r = phase_row(phase, ri);
        const int rn = ri + FR_WARPS;
        if (rn < nrows) {
            const uint16_t* rp = base + (size_t)phase_row(phase, rn) * vw.pitch;
#pragma unroll
            for (int k = 0; k < FR_VPL; k++) qn[k] = full[k] ? ldg_stream16(rp + (lane + 32 * k) * 8) : make_uint4(0, 0, 0, 0);
        }
        uint32_t rs = 0;
#pragma unroll
        for (int k = 0; k < FR_VPL; k++) {
            if (full[k]) {
                const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
                const uint32_t vm = __vminu2(__vminu2(w[0], w[1]), __vminu2(w[2], w[3]));
                const uint32_t vM = __vmaxu2(__vmaxu2(w[0], w[1]), __vmaxu2(w[2], w[3]));
                acc.mn2 = __vminu2(acc.mn2, vm);
                acc.mx2 = __vmaxu2(acc.mx2, vM);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    acc.cs[k * 8 + 2 * t] = dp2a(w[t], 0x0001u, acc.cs[k * 8 + 2 * t]);
                    acc.cs[k * 8 + 2 * t + 1] = dp2a(w[t], 0x0100u, acc.cs[k * 8 + 2 * t + 1]);
                    rs = dp2a(w[t], 0x0101u, rs);
                    band_px(w[t] & 0xffffu, bd[1], hist + FR_CAP, acc.below[1], acc.le[1], MED_E);
                    band_px(w[t] >> 16, bd[1], hist + FR_CAP, acc.below[1], acc.le[1], MED_E);
                }
                if (lo_on && (lo_always || __vminu2(vm, lo_c) != lo_c)) {
                    const uint2 d = band_vec(q[k], bd[0].a, bd[0].acc_n, bd[0].off, bd[0].e, 0u, hist);
                    acc.below[0] += d.x;
                    acc.le[0] += d.y;
                }
                if (hi_on && (hi_always || __vmaxu2(vM, hi_c) != hi_c)) {
                    const uint2 d = band_vec(q[k], bd[2].a, bd[2].acc_n, bd[2].off, bd[2].e, 1u, hist + 2 * FR_CAP);
                    acc.below[2] += d.x;
                    acc.le[2] += d.y;
                }
            }
        }
        rs = __reduce_add_sync(0xffffffffu, rs);
        if (lane == 0) rowsum_sm[r] += rs;
#pragma unroll
        for (int k = 0; k < FR_VPL; k++) q[k] = qn[k];
        ri = rn;
    }
}

// Columns of a misaligned view that no full vector covers (< 8 on each side): scalar pass, counted as "main" pixels.
__device__ inline void edge_columns(const View& vw, const BandRT (&bd)[FR_NB], bool lo_on, bool hi_on, uint32_t* __restrict__ hist,
                                    uint32_t* __restrict__ rowsum_sm, uint32_t* __restrict__ colsum_sm, Acc& acc) {
    // full vectors cover view columns [cl, cr)
    const int jf = (vw.mis + 7) / 8;                 // first vector with c0 >= 0
    const int jl = (vw.W + vw.mis) / 8;              // vectors j < jl have c0 + 8 <= W
    const int cl = min(vw.W, max(0, jf * 8 - vw.mis)), cr = jl > jf ? jl * 8 - vw.mis : cl;
    const int ne = cl + (vw.W - cr);
    if (ne <= 0) return;
    uint32_t mnv = 0xffffu, mxv = 0;
    for (int i = threadIdx.x; i < ne * vw.H; i += FR_THREADS) {
        const int r = i / ne, e = i - r * ne;
        const int cidx = e < cl ? e : cr + (e - cl);
        const uint32_t v = __ldg(vw.f + (size_t)r * vw.pitch + cidx);
        mnv = min(mnv, v);
        mxv = max(mxv, v);
        atomicAdd(&rowsum_sm[r], v);
        atomicAdd(&colsum_sm[cidx], v);
        band_px(v, bd[1], hist + FR_CAP, acc.below[1], acc.le[1], bd[1].e != 0);
        if (lo_on) band_px(v, bd[0], hist, acc.below[0], acc.le[0], bd[0].e != 0);
        if (hi_on) band_px(65535u - v, bd[2], hist + 2 * FR_CAP, acc.below[2], acc.le[2], bd[2].e != 0);
    }
    acc.mn2 = __vminu2(acc.mn2, mnv * 0x00010001u);
    acc.mx2 = __vmaxu2(acc.mx2, mxv * 0x00010001u);
}

// registers -> shared column sums (accumulating), two rounds of 8 warps through `stage` (FR_STAGE_WARPS * 1024 words)
__device__ inline void flush_colsum(uint32_t (&cs)[FR_VPL * 8], uint32_t* __restrict__ stage, uint32_t* __restrict__ colsum_sm, const View& vw, int strip) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int round = 0; round < FR_WARPS / FR_STAGE_WARPS; round++) {
        __syncthreads();
        if (wid / FR_STAGE_WARPS == round) {
            uint32_t* st = stage + (wid % FR_STAGE_WARPS) * (FR_STRIP_VEC * 8);
#pragma unroll
            for (int kp = 0; kp < FR_VPL * 8; kp++) st[kp * 32 + lane] = cs[kp];
        }
        __syncthreads();
        // thread -> (l = tid & 31, kp = tid >> 5 [+16])
        for (int kp = tid >> 5; kp < FR_VPL * 8; kp += FR_WARPS) {
            uint32_t s = 0;
#pragma unroll
            for (int w8 = 0; w8 < FR_STAGE_WARPS; w8++) s += stage[w8 * (FR_STRIP_VEC * 8) + kp * 32 + lane];
            const int k = kp >> 3, p = kp & 7;
            const int j = strip * FR_STRIP_VEC + lane + 32 * k;
            const int c = j * 8 - vw.mis + p;
            if (c >= 0 && c < vw.W && s) colsum_sm[c] += s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int kp = 0; kp < FR_VPL * 8; kp++) cs[kp] = 0;
}

// merge the per-thread band counters of one phase into shared memory
__device__ inline void merge_band_counters(Acc& acc, const BandRT (&bd)[FR_NB], const uint32_t (&on)[FR_NB], uint32_t* __restrict__ hist,
                                           uint32_t* __restrict__ below_out) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < FR_NB; j++) {
        uint32_t b = __reduce_add_sync(0xffffffffu, acc.below[j]);
        uint32_t l = __reduce_add_sync(0xffffffffu, acc.le[j]);
        if (lane == 0 && on[j]) {
            if (b) atomicAdd(&below_out[j], b);
            if (bd[j].e && l > b) atomicAdd(&hist[j * FR_CAP + bd[j].off], l - b);   // pixels equal to the lower edge
        }
        acc.below[j] = 0;
        acc.le[j] = 0;
    }
}

__global__ void __launch_bounds__(FR_THREADS, 1)
k_pf_front(const PfConst* __restrict__ cc, const StatsGeom g, const FrameRef* __restrict__ frames, int nframes, PfFrame* fr,
           FrameStats* __restrict__ stats, int* counters) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const PfConst& c = *cc;
    const int H = g.H, W = g.W;
    const int Hp = (H + 3) & ~3, Wp = (W + 3) & ~3;
    // layout: [FrontSh][rowsum Hp][colsum Wp][rowsum2 Hp][colsum2 Wp][work: hist 3*CAP + stage 8*1024 | profile scratch]
    FrontSh* sh = reinterpret_cast<FrontSh*>(smraw);
    uint32_t* rowsum_sm = reinterpret_cast<uint32_t*>(smraw + ((sizeof(FrontSh) + 15) & ~(size_t)15));
    uint32_t* colsum_sm = rowsum_sm + Hp;
    uint32_t* rowsum2_sm = colsum_sm + Wp;
    uint32_t* colsum2_sm = rowsum2_sm + Hp;
    uint32_t* hist = colsum2_sm + Wp;
    uint32_t* stage = hist + FR_NB * FR_CAP;
    unsigned char* prof_raw = reinterpret_cast<unsigned char*>(hist);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t npix = (uint32_t)H * (uint32_t)W;
    // target ranks (pairs): [0] low pair, [1] median pair, [2] high pair
    const uint32_t rk[FR_NB][2] = {{g.ranks[0], g.ranks[1]}, {g.ranks[4], g.ranks[5]}, {g.ranks[2], g.ranks[3]}};

    for (int fi = blockIdx.x; fi < nframes; fi += gridDim.x) {
        PfFrame& f = fr[fi];
        const FrameRef frf = frames[fi];
        View vw;
        vw.f = frf.origin;
        vw.pitch = frf.pitch;
        vw.H = H;
        vw.W = W;
        vw.mis = (int)((reinterpret_cast<uintptr_t>(frf.origin) >> 1) & 7);
        vw.nvec = (W + vw.mis + 7) / 8;
        const int nstrips = (vw.nvec + FR_STRIP_VEC - 1) / FR_STRIP_VEC;
        // ---- reset
        for (int i = tid; i < FR_NB * FR_CAP; i += FR_THREADS) hist[i] = 0;
        for (int i = tid; i < W; i += FR_THREADS) colsum_sm[i] = 0;
        for (int i = tid; i < H; i += FR_THREADS) rowsum_sm[i] = 0;
        if (tid < FR_NB) { sh->below_p[tid] = 0; sh->below_m[tid] = 0; }
        if (tid == 0) { sh->fallback = 0; sh->corner = 0; sh->mn = 0xffffu; sh->mx = 0; sh->pmn = 0xffffu; sh->pmx = 0; }
        __syncthreads();

        // ---- T0: 2048-pixel sample (8 rows x 64 threads x 4 px), value bisection for the wide band edges
        uint32_t sv[FR_T0];
        {
            const int srow = min(H - 1, (int)(((2 * (tid >> 6) + 1) * (long long)H) / 16));
            const int seg = tid & 63;
            const int scol = min(max(W - FR_T0, 0), (int)(((long long)seg * W) / 64));
#pragma unroll
            for (int p = 0; p < FR_T0; p++) sv[p] = __ldg(vw.f + (size_t)srow * vw.pitch + min(scol + p, W - 1));
        }
        const uint32_t nT0 = FR_THREADS * FR_T0;
        if (tid < 8) {
            // targets: t0: band0 upper, t1: band1 lower, t2: band1 upper, t3: band2 upper (flipped domain)
            // sample rank of full rank k: k * nT0 / npix; margins 5 sigma + 2
            double fr_[3];
            fr_[0] = (double)rk[0][1] / (double)npix;
            fr_[1] = (double)rk[1][0] / (double)npix;
            fr_[2] = (double)(npix - 1 - rk[2][0]) / (double)npix;
            uint32_t rank = 0;
            if (tid < 4) {
                const int b = tid == 0 ? 0 : (tid == 3 ? 2 : 1);
                const double fq = fr_[b];
                const double sg = sqrt(fq * (1.0 - fq) * (double)nT0);
                const double ctr = fq * (double)nT0;
                double rr = (tid == 1) ? ctr - 5.0 * sg - 2.0 : ctr + 5.0 * sg + 3.0;
                rr = fmin(fmax(rr, 0.0), (double)(nT0 - 1));
                rank = (uint32_t)rr;
            }
            sh->bis_rank[tid] = rank;
            sh->bis_lo[tid] = 0;
            sh->bis_hi[tid] = 65535u;
            sh->bis_cnt[tid] = 0;
        }
        {   // sample min / max
            uint32_t mnv = min(min(sv[0], sv[1]), min(sv[2], sv[3])), mxv = max(max(sv[0], sv[1]), max(sv[2], sv[3]));
            mnv = warp_min(mnv);
            mxv = warp_max(mxv);
            if (lane == 0) { atomicMin(&sh->pmn, mnv); atomicMax(&sh->pmx, mxv); }
        }
        __syncthreads();
        for (int step = 0; step < 16; step++) {
            // count(v' <= mid) for the 4 targets
            uint32_t cnt[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint32_t mid = (sh->bis_lo[t] + sh->bis_hi[t]) >> 1;
                uint32_t cc_ = 0;
#pragma unroll
                for (int p = 0; p < FR_T0; p++) {
                    const uint32_t v = t == 3 ? 65535u - sv[p] : sv[p];
                    cc_ += v <= mid ? 1u : 0u;
                }
                cnt[t] = __reduce_add_sync(0xffffffffu, cc_);
            }
            if (lane < 4) atomicAdd(&sh->bis_cnt[lane], lane == 0 ? cnt[0] : lane == 1 ? cnt[1] : lane == 2 ? cnt[2] : cnt[3]);
            __syncthreads();
            if (tid < 4) {
                const uint32_t mid = (sh->bis_lo[tid] + sh->bis_hi[tid]) >> 1;
                if (sh->bis_cnt[tid] >= sh->bis_rank[tid] + 1) sh->bis_hi[tid] = mid; else sh->bis_lo[tid] = mid + 1;
                sh->bis_cnt[tid] = 0;
            }
            __syncthreads();
        }
        if (tid == 0) {
            const uint32_t smin = sh->pmn, smax = sh->pmx;
            // band 0: [sample min, q(upper)], band 1: [q(lower), q(upper)], band 2 (flipped): [65535 - sample max, q'(upper)]
            uint32_t a[FR_NB], b[FR_NB];
            a[0] = smin;              b[0] = max(sh->bis_lo[0], smin);
            a[1] = sh->bis_lo[1];     b[1] = max(sh->bis_lo[2], a[1]);
            a[2] = 65535u - smax;     b[2] = max(sh->bis_lo[3], a[2]);
            for (int j = 0; j < FR_NB; j++) {
                if (b[j] - a[j] > FR_CAP - 1) {
                    // too wide for the histogram: keep the part next to the expected target
                    if (j == 1) { const uint32_t mid = a[j] + (b[j] - a[j]) / 2; a[j] = mid - (FR_CAP / 2 - 1); b[j] = a[j] + FR_CAP - 1; }
                    else b[j] = a[j] + FR_CAP - 1;
                }
                sh->a0[j] = a[j];
                sh->b0[j] = b[j];
            }
            sh->pmn = 0xffffu;
            sh->pmx = 0;
        }
        __syncthreads();

        // ---- P: pilot rows with the wide bands (all on, lower edges counted in registers)
        Acc acc;
#pragma unroll
        for (int i = 0; i < FR_VPL * 8; i++) acc.cs[i] = 0;
        acc.mn2 = 0xffffffffu;
        acc.mx2 = 0;
#pragma unroll
        for (int j = 0; j < FR_NB; j++) { acc.below[j] = 0; acc.le[j] = 0; }
        BandRT bd[FR_NB];
#pragma unroll
        for (int j = 0; j < FR_NB; j++) {
            bd[j].a = sh->a0[j];
            bd[j].e = 1;
            bd[j].acc_n = sh->b0[j] - sh->a0[j];       // accepted: a+1 .. b
            bd[j].off = 0;
        }
        const uint32_t on_all[FR_NB] = {1, 1, 1};
        for (int s = 0; s < nstrips; s++) {
            stream_phase<true>(vw, s, 0, bd, true, true, hist, rowsum_sm, acc);
            if (nstrips > 1) flush_colsum(acc.cs, stage, colsum_sm, vw, s);
        }
        merge_band_counters(acc, bd, on_all, hist, sh->below_p);
        {
            uint32_t mnv = min(acc.mn2 & 0xffffu, acc.mn2 >> 16), mxv = max(acc.mx2 & 0xffffu, acc.mx2 >> 16);
            mnv = warp_min(mnv);
            mxv = warp_max(mxv);
            if (lane == 0) { atomicMin(&sh->pmn, mnv); atomicMax(&sh->pmx, mxv); }
        }
        __syncthreads();

        // ---- R: re-pivot from the pilot histogram
        const int e_jf = (vw.mis + 7) / 8, e_jl = (W + vw.mis) / 8;
        const int e_cl = min(W, max(0, e_jf * 8 - vw.mis)), e_cr = e_jl > e_jf ? e_jl * 8 - vw.mis : e_cl;   // full vectors cover [cl, cr)
        const uint32_t n_p = (uint32_t)pilot_rows(H) * (uint32_t)(e_cr - e_cl);
        for (int j = 0; j < FR_NB; j++) {
            const uint32_t a0 = sh->a0[j], nb = sh->b0[j] - a0 + 1;
            // target ranks in the band's domain
            const uint32_t k_lo = j == 2 ? npix - 1 - rk[2][1] : rk[j][0];
            const uint32_t k_hi = j == 2 ? npix - 1 - rk[2][0] : rk[j][1];
            uint32_t fa = 0, fb = nb - 1;
            if (n_p > 0) {
                const double scale = (double)n_p / (double)npix;
                const double fq = ((double)k_lo + 0.5) / (double)npix;
                const double sg = sqrt(fq * (1.0 - fq) * (double)n_p);
                const double rl = (double)k_lo * scale - 5.0 * sg - 2.0;
                const double ru = (double)k_hi * scale + 5.0 * sg + 3.0;
                const uint32_t base = sh->below_p[j];
                const uint32_t r_l = rl <= 0.0 ? 0u : (uint32_t)rl;
                const uint32_t r_u = (uint32_t)fmin(ru, (double)n_p);
                const uint32_t xa = hist_find(hist + j * FR_CAP, nb, base, r_l, sh, 0);
                const uint32_t xb = hist_find(hist + j * FR_CAP, nb, base, r_u, sh, 1);
                fa = xa == 0xffffffffu ? 0u : min(xa, nb - 1);   // rank below the band: keep the pilot edge
                fb = xb == 0xffffffffu ? 0u : min(xb, nb - 1);
                if (fb < fa) fb = fa;
            }
            if (tid == 0) {
                sh->a1[j] = a0 + fa;
                sh->b1[j] = a0 + fb;
                const uint32_t cnt_a = hist[j * FR_CAP + fa];
                sh->e1[j] = (n_p > 0 && cnt_a >= max(16u, n_p >> 10)) ? 1u : 0u;
                sh->on1[j] = 1;
                sh->cand[j] = 0;
                // switch-off proof for the extreme bands: the pilot already holds k_hi + 1 pixels equal to its minimum
                if (j != 1 && n_p > 0 && sh->below_p[j] == 0) {
                    const uint32_t pm = j == 0 ? sh->pmn : 65535u - sh->pmx;   // pilot minimum in the band's domain
                    if (pm >= a0 && pm - a0 < nb && hist[j * FR_CAP + (pm - a0)] >= k_hi + 1) { sh->on1[j] = 0; sh->cand[j] = pm; }
                }
            }
            __syncthreads();
        }

        // ---- M: the remaining rows with the narrow bands
        uint32_t on_m[FR_NB];
#pragma unroll
        for (int j = 0; j < FR_NB; j++) {
            bd[j].a = sh->a1[j];
            bd[j].e = sh->e1[j];
            bd[j].acc_n = sh->b1[j] - sh->a1[j] + 1 - sh->e1[j];
            bd[j].off = sh->a1[j] - sh->a0[j];
            on_m[j] = sh->on1[j];
        }
        for (int s = 0; s < nstrips; s++) {
            if (bd[1].e) stream_phase<true>(vw, s, 1, bd, on_m[0] != 0, on_m[2] != 0, hist, rowsum_sm, acc);
            else stream_phase<false>(vw, s, 1, bd, on_m[0] != 0, on_m[2] != 0, hist, rowsum_sm, acc);
            flush_colsum(acc.cs, stage, colsum_sm, vw, s);
        }
        edge_columns(vw, bd, on_m[0] != 0, on_m[2] != 0, hist, rowsum_sm, colsum_sm, acc);
        merge_band_counters(acc, bd, on_m, hist, sh->below_m);
        {
            uint32_t mnv = min(acc.mn2 & 0xffffu, acc.mn2 >> 16), mxv = max(acc.mx2 & 0xffffu, acc.mx2 >> 16);
            mnv = warp_min(mnv);
            mxv = warp_max(mxv);
            if (lane == 0) { atomicMin(&sh->mn, mnv); atomicMax(&sh->mx, mxv); }
        }
        // corner boxes (core/image.py:881-894)
        if (g.box > 0) {
            const int per = g.box * g.box;
            unsigned long long cs = 0;
            for (int i = tid; i < 4 * per; i += FR_THREADS) {
                const int b = i / per, o = i - b * per;
                const int y = o / g.box, x = o - y * g.box;
                const int rr = ((b & 2) ? H - g.rp - g.box : g.rp) + y;
                const int cl = ((b & 1) ? W - g.cp - g.box : g.cp) + x;
                if (rr >= 0 && rr < H && cl >= 0 && cl < W) cs += __ldg(vw.f + (size_t)rr * vw.pitch + cl);
            }
            cs = warp_sum(cs);
            if (lane == 0 && cs) atomicAdd(&sh->corner, cs);
        }
        __syncthreads();
        if (tid == 0) { sh->mn = min(sh->mn, sh->pmn); sh->mx = max(sh->mx, sh->pmx); }
        // total sum = sum of the row sums
        unsigned long long tsum = 0;
        for (int i = tid; i < H; i += FR_THREADS) tsum += rowsum_sm[i];
        tsum = warp_sum(tsum);
        __syncthreads();
        if (tid == 0) sh->st.sum = 0;
        __syncthreads();
        if (lane == 0) atomicAdd(&sh->st.sum, tsum);

        // ---- X: resolve the order statistics
        for (int j = 0; j < FR_NB; j++) {
            const int s0 = j == 0 ? 0 : (j == 1 ? 4 : 2);      // FrameStats.ostat slot of the pair's lower rank
            if (!sh->on1[j]) {
                if (tid == 0) {
                    const uint32_t gm = j == 0 ? sh->mn : 65535u - sh->mx;   // global minimum in the band's domain
                    if (gm != sh->cand[j]) sh->fallback = 1;
                    const uint32_t raw = j == 2 ? 65535u - sh->cand[j] : sh->cand[j];
                    sh->st.ostat[s0] = raw;
                    sh->st.ostat[s0 + 1] = raw;
                }
                __syncthreads();
                continue;
            }
            const uint32_t a0 = sh->a0[j], a1 = sh->a1[j], b1 = sh->b1[j];
            const uint32_t pre = hist_sum_below(hist + j * FR_CAP, a1 - a0, sh);      // pilot pixels in [a0, a1)
            const uint32_t base = sh->below_p[j] + pre + sh->below_m[j];               // pixels below a1 (whole frame)
            const uint32_t k_lo = j == 2 ? npix - 1 - rk[2][1] : rk[j][0];
            const uint32_t k_hi = j == 2 ? npix - 1 - rk[2][0] : rk[j][1];
            const uint32_t nb = b1 - a1 + 1;
            const uint32_t x_lo = hist_find(hist + j * FR_CAP + (a1 - a0), nb, base, k_lo, sh, 2);
            const uint32_t x_hi = hist_find(hist + j * FR_CAP + (a1 - a0), nb, base, k_hi, sh, 3);
            if (tid == 0) {
                if (x_lo >= nb || x_hi >= nb) sh->fallback = 1;
                const uint32_t v_lo = a1 + min(x_lo, nb - 1), v_hi = a1 + min(x_hi, nb - 1);
                if (j == 2) { sh->st.ostat[s0] = 65535u - v_hi; sh->st.ostat[s0 + 1] = 65535u - v_lo; }
                else { sh->st.ostat[s0] = v_lo; sh->st.ostat[s0 + 1] = v_hi; }
            }
            __syncthreads();
        }
        // ---- decide (noise, inversion, median in g units)
        if (tid == 0) {
            sh->st.mn = sh->mn;
            sh->st.mx = sh->mx;
            sh->st.npix = npix;
            sh->st.corner_sum = sh->corner;
            sh->st.overflow = sh->fallback ? 1u : 0u;
            stats[fi] = sh->st;
            f.status = EPID_PF_OK;
            f.noisy = 0;
            f.inv = 0;
            f.corner_inverted = 0;
            f.noise_passes = 0;
            f.n_pickets = 0;
            f.n_inview = 0;
            f.todo = 0;
            f.orientation = 0;
            if (sh->fallback) atomicAdd(&counters[1], 1);
            pf_decide_frame(c, sh->st, f, 1, counters);
        }
        __syncthreads();
        if (sh->fallback || f.status != EPID_PF_OK) { __syncthreads(); continue; }

        // ---- clamped sums: sum of max(2g, med2) per row / column (picketfence.py:1509-1514), second sweep (L2 / HBM)
        if (c.p.orientation < 0) {
            const uint32_t inv = f.inv, mn = f.mn, mx = f.mx, med2 = f.med2;
            // raw-domain clamp constant C and parity handling (see DESIGN.md "clamped sums")
            uint32_t C, odd;
            if (!inv) { const uint32_t t2 = 2u * mn + med2; odd = t2 & 1u; C = (t2 + 1u) >> 1; }
            else { const uint32_t u2 = 2u * mx - med2; odd = u2 & 1u; C = (u2 >> 1) + odd; if (C > 65535u) { C = 65535u; odd = 0; } }
            const uint32_t Cp = C | (C << 16);
            uint32_t* fsum = stage + FR_STAGE_WARPS * FR_STRIP_VEC * 8;   // [W] per-column flag counts (only when odd)
            for (int i = tid; i < W; i += FR_THREADS) colsum2_sm[i] = 0;
            for (int i = tid; i < W + H; i += FR_THREADS) fsum[i] = 0;
            __syncthreads();
            for (int i = tid; i < H; i += FR_THREADS) rowsum2_sm[i] = 0;     // holds sum(x) - (flag count) per row first
            __syncthreads();
            uint32_t fsp[FR_VPL * 4];    // packed u16x2 per-column flag counts (only used when odd; <= H / FR_WARPS per thread)
            for (int s = 0; s < nstrips; s++) {
#pragma unroll
                for (int i = 0; i < FR_VPL * 8; i++) acc.cs[i] = 0;
#pragma unroll
                for (int i = 0; i < FR_VPL * 4; i++) fsp[i] = 0;
                bool full[FR_VPL];
#pragma unroll
                for (int k = 0; k < FR_VPL; k++) {
                    const int j = s * FR_STRIP_VEC + lane + 32 * k;
                    const int c0 = j * 8 - vw.mis;
                    full[k] = j < vw.nvec && c0 >= 0 && c0 + 8 <= W;
                }
                const uint16_t* __restrict__ base = vw.f - vw.mis + (size_t)s * FR_STRIP_VEC * 8;
                for (int r = wid; r < H; r += FR_WARPS) {
                    const uint16_t* rp = base + (size_t)r * vw.pitch;
                    uint4 q[FR_VPL];
#pragma unroll
                    for (int k = 0; k < FR_VPL; k++) q[k] = full[k] ? ldg_stream16(rp + (lane + 32 * k) * 8) : make_uint4(0, 0, 0, 0);
                    uint32_t rs = 0, rf = 0;
#pragma unroll
                    for (int k = 0; k < FR_VPL; k++) {
                        if (full[k]) {
                            const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                const uint32_t x = inv ? __vminu2(w[t], Cp) : __vmaxu2(w[t], Cp);
                                acc.cs[k * 8 + 2 * t] = dp2a(x, 0x0001u, acc.cs[k * 8 + 2 * t]);
                                acc.cs[k * 8 + 2 * t + 1] = dp2a(x, 0x0100u, acc.cs[k * 8 + 2 * t + 1]);
                                rs = dp2a(x, 0x0101u, rs);
                                if (odd) {
                                    // !inv: [v < C] = [max(v,C) - v > 0];  inv: [v < C] = [C - min(v,C) > 0]  (no borrow between halves)
                                    const uint32_t d = inv ? Cp - x : x - w[t];
                                    const uint32_t fl = __vminu2(d, 0x00010001u);
                                    fsp[k * 4 + t] += fl;
                                    rf = dp2a(fl, 0x0101u, rf);
                                }
                            }
                        }
                    }
                    rs = __reduce_add_sync(0xffffffffu, rs);
                    rf = __reduce_add_sync(0xffffffffu, rf);
                    if (lane == 0) { rowsum2_sm[r] += rs; if (odd) atomicAdd(&fsum[W + r], rf); }
                }
                // column partials: colsum2_sm += sum(x); fsum += flag counts (only when odd)
                flush_colsum(acc.cs, stage, colsum2_sm, vw, s);
                if (odd) {
                    uint32_t fs[FR_VPL * 8];
#pragma unroll
                    for (int i = 0; i < FR_VPL * 4; i++) { fs[2 * i] = fsp[i] & 0xffffu; fs[2 * i + 1] = fsp[i] >> 16; }
                    flush_colsum(fs, stage, fsum, vw, s);
                }
            }
            // edge columns of a misaligned view
            {
                const int ne = e_cl + (W - e_cr);
                for (int i = tid; i < ne * H; i += FR_THREADS) {
                    const int r = i / ne, e = i - r * ne;
                    const int cidx = e < e_cl ? e : e_cr + (e - e_cl);
                    const uint32_t v = __ldg(vw.f + (size_t)r * vw.pitch + cidx);
                    const uint32_t x = inv ? min(v, C) : max(v, C);
                    atomicAdd(&rowsum2_sm[r], x);
                    atomicAdd(&colsum2_sm[cidx], x);
                    if (odd && v < C) { atomicAdd(&fsum[W + r], 1u); atomicAdd(&fsum[cidx], 1u); }
                }
            }
            __syncthreads();
            for (int y = tid; y < H; y += FR_THREADS) {
                const uint32_t sx = rowsum2_sm[y];
                const uint32_t ff = odd ? fsum[W + y] : 0u;
                rowsum2_sm[y] = !inv ? 2u * sx - ff - 2u * mn * (uint32_t)W : 2u * mx * (uint32_t)W - 2u * sx + (odd ? (uint32_t)W - ff : 0u);
            }
            for (int x = tid; x < W; x += FR_THREADS) {
                const uint32_t sx = colsum2_sm[x];
                const uint32_t ff = odd ? fsum[x] : 0u;
                colsum2_sm[x] = !inv ? 2u * sx - ff - 2u * mn * (uint32_t)H : 2u * mx * (uint32_t)H - 2u * sx + (odd ? (uint32_t)H - ff : 0u);
            }
            __syncthreads();
        }
        // ---- orientation, leaf profile, pickets (pf_profile_block aliases the histogram / stage area)
        __syncthreads();
        pf_profile_block(c, f, rowsum_sm, colsum_sm, rowsum2_sm, colsum2_sm, prof_raw);
        __syncthreads();
    }
}

size_t pf_front_smem_bytes(int H, int W) {
    const int Hp = (H + 3) & ~3, Wp = (W + 3) & ~3;
    size_t work = sizeof(uint32_t) * (size_t)(FR_NB * FR_CAP + FR_STAGE_WARPS * FR_STRIP_VEC * 8 + Wp + Hp + 64);
    const size_t prof = pf_profile_smem_bytes(FR_THREADS);
    if (prof > work) work = prof;
    return ((sizeof(FrontSh) + 15) & ~(size_t)15) + sizeof(uint32_t) * (size_t)(2 * Hp + 2 * Wp) + work + 64;
}

bool pf_front_supported(int H, int W, int pitch) {
    return (pitch % 8) == 0 && H >= 64 && W >= 64 && H <= STATS_MAX_DIM && W <= STATS_MAX_DIM && pf_front_smem_bytes(H, W) <= 220 * 1024;
}

int launch_pf_front(epid_ctx* ctx, cudaStream_t stream, const PfConst* d_cst, const StatsGeom& g, const FrameRef* refs, int n, PfFrame* fr,
                    FrameStats* stats, int* counters) {
    static size_t attr = 0;
    const size_t smem = pf_front_smem_bytes(g.H, g.W);
    if (smem > attr) {
        EPID_CUDA(cudaFuncSetAttribute(k_pf_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const int grid = n < ctx->sm_count ? n : ctx->sm_count;
    k_pf_front<<<grid, FR_THREADS, smem, stream>>>(d_cst, g, refs, n, fr, stats, counters);
    ctx->launches++;
    EPID_CUDA(cudaGetLastError());
    return EPID_OK;
}

}  // namespace epid
