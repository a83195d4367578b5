// Fast per-(leaf, picket) window kernel of the PicketFence pipeline.
//
// Reference semantics: PicketFence._get_mlc_window / _is_mlc_peak_in_window (picketfence.py:847-886) and
// MLCValue.get_peak_positions (picketfence.py:1605-1628) -> FWXMProfilePhysical.field_edge_idx (core/profile.py:602-611).
//
// One warp per window, 8 windows in flight per CTA, the windows of a frame spread over gridDim.x CTAs.
//   1. stage the window as exact integers g (ground / invert folded in) into shared memory, canonical layout
//      px[i * S + j]: i across the leaf (the axis np.median collapses), j along leaf travel; the row stride S is even
//      with S/2 odd, so "lane = row" accesses are bank-conflict free and a 32-bit load yields two adjacent travel samples;
//   2. validity: lanes own rows -> sum / sum of squares / max along travel with no shuffles (exact integer variance);
//   3. np.median(window, axis): lanes own PAIRS of travel samples held packed u16x2 in registers and sorted by a fully
//      unrolled Batcher odd-even merge network on VIMNMX.U16x2 (sm_100a packed 16-bit min/max);
//   4. the 1-D profile (<= 256 samples) is normalised in fp64 and searched for its most prominent peak, FWHM edges by
//      scipy's peak_widths interpolation -- same operation order as the reference.
// Windows that do not fit the fast path (nc > 256, nr > 64 or nr * S > W2_CAP) are marked valid = -1 and picked up by the
// generic kernel (k_pf_windows in pf.cu) launched right after in "todo" mode.
#include "pf_common.cuh"

namespace epid {

constexpr int W2_WARPS = 8;
constexpr int W2_CAP = 6144;     // staged u16 elements per warp
constexpr int W2_MAXNC = 256;    // travel samples per window on the fast path
constexpr int W2_GRID_X = 64;    // CTAs per frame

template <int N>
__device__ __forceinline__ void sort_net_u16x2(uint32_t (&r)[N]) {
    // Batcher odd-even merge sort, ascending in both 16-bit halves independently
#pragma unroll
    for (int p = 1; p < N; p <<= 1) {
#pragma unroll
        for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
            for (int j = k % p; j <= N - 1 - k; j += 2 * k) {
#pragma unroll
                for (int i = 0; i < k; i++) {
                    if (i <= N - j - k - 1 && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        const uint32_t a = r[i + j], b = r[i + j + k];
                        r[i + j] = __vminu2(a, b);
                        r[i + j + k] = __vmaxu2(a, b);
                    }
                }
            }
        }
    }
}

// 2 * median over the nr rows of the travel-sample pair `t` (columns 2t, 2t+1): returns (va + vb) per half
template <int NRP>
__device__ __forceinline__ void pair_median2(const uint16_t* __restrict__ px, int S, int nr, int t, int k1, int k2,
                                             uint32_t& m_lo, uint32_t& m_hi) {
    uint32_t r[NRP];
#pragma unroll
    for (int i = 0; i < NRP; i++)
        r[i] = i < nr ? *reinterpret_cast<const uint32_t*>(px + i * S + 2 * t) : 0xffffffffu;
    sort_net_u16x2<NRP>(r);
    uint32_t va = 0, vb = 0;
#pragma unroll
    for (int i = 0; i < NRP; i++) {
        if (i == k1) va = r[i];
        if (i == k2) vb = r[i];
    }
    m_lo = (va & 0xffffu) + (vb & 0xffffu);
    m_hi = (va >> 16) + (vb >> 16);
}

__global__ void __launch_bounds__(W2_WARPS * 32)
k_pf_windows_fast(const PfConst* __restrict__ cc, const FrameRef* __restrict__ frames, PfFrame* fr, PfWin* __restrict__ wins) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int fi = blockIdx.y;
    const PfConst& c = *cc;
    PfFrame& f = fr[fi];
    if (f.status != EPID_PF_OK) return;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int np = f.n_pickets;
    const int total = f.n_inview * np;
    uint16_t* px = reinterpret_cast<uint16_t*>(smraw) + (size_t)wid * W2_CAP;
    uint32_t* m2 = reinterpret_cast<uint32_t*>(smraw + sizeof(uint16_t) * W2_CAP * W2_WARPS) + (size_t)wid * W2_MAXNC;
    double* xs = reinterpret_cast<double*>(px);
    const int H = c.H, W = c.W;
    const int orient = f.orientation;
    const double dpmm = c.p.dpmm;
    const FrameRef frf = frames[fi];
    const int inv = f.inv;
    const uint32_t mn = f.mn, mx = f.mx;
    const double Dd = (double)f.D;
    const double spacing = f.spacing;
    const int sag = c.p.sag_px;

    for (int widx = blockIdx.x * W2_WARPS + wid; widx < total; widx += gridDim.x * W2_WARPS) {
        const int li = widx / np, pk = widx - li * np;
        const int leaf = f.inview[li];
        const double lw_px = c.p.leaf_width_mm[leaf] * dpmm;
        const double lc_px = c.p.leaf_center_mm[leaf] * dpmm + (orient == 0 ? (double)H / 2.0 : (double)W / 2.0);
        PfWin& out = wins[((size_t)fi * PF_L + li) * PF_P + pk];
        const double pidx = (double)f.picket_idx[pk];
        // _get_mlc_window (picketfence.py:859-886): python int() truncates toward zero
        const int a0 = max((int)(pidx - spacing / 2.0), 0);                                   // along travel
        const int a1 = min((int)(pidx + spacing / 2.0), orient == 0 ? W : H);
        const int b0 = max((int)(lc_px - lw_px / 2.0), 0);                                    // across the leaf
        const int b1 = min((int)(lc_px + lw_px / 2.0), orient == 0 ? H : W);
        const int nc = a1 - a0, nr = b1 - b0;
        if (nc <= 0 || nr <= 0) {           // empty slice: np.max raises ValueError in the reference
            if (lane == 0) { out.valid = 0; out.l = 0; out.r = 0; f.status = EPID_PF_WINDOW_NO_PEAK; }
            continue;
        }
        int S = (nc + 1) & ~1;
        if (((S >> 1) & 1) == 0) S += 2;
        if (nc > W2_MAXNC || nr > 64 || nr * S > W2_CAP) {
            if (lane == 0) out.valid = -1;   // generic kernel
            continue;
        }
        __syncwarp();
        // ---- 1. stage (np.roll(sag) folded into the source index)
        uint32_t gmax = 0;
        if (orient == 0) {
            for (int i = 0; i < nr; i++) {
                int row = b0 + i - sag;
                if (sag) { row %= H; if (row < 0) row += H; }
                const uint16_t* __restrict__ src = frf.origin + (size_t)row * frf.pitch + a0;
                for (int jj = lane; jj < S; jj += 32) {
                    uint32_t g = 0;
                    if (jj < nc) {
                        const uint32_t v = __ldg(src + jj);
                        g = inv ? mx - v : v - mn;
                        gmax = max(gmax, g);
                    }
                    px[i * S + jj] = (uint16_t)g;
                }
            }
        } else {
            for (int jj = 0; jj < nc; jj++) {
                const uint16_t* __restrict__ src = frf.origin + (size_t)(a0 + jj) * frf.pitch;
                for (int i = lane; i < nr; i += 32) {
                    int col = b0 + i - sag;
                    if (sag) { col %= W; if (col < 0) col += W; }
                    const uint32_t v = __ldg(src + col);
                    const uint32_t g = inv ? mx - v : v - mn;
                    gmax = max(gmax, g);
                    px[i * S + jj] = (uint16_t)g;
                }
            }
            for (int i = lane; i < nr; i += 32)
                for (int jj = nc; jj < S; jj++) px[i * S + jj] = 0;
        }
        gmax = warp_max(gmax);
        __syncwarp();
        // ---- 2. _is_mlc_peak_in_window (picketfence.py:847-857): lanes own rows
        const int k1 = (nr - 1) / 2, k2 = nr / 2;
        double sd[2] = {-1.0, -1.0};
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            const int i = sl * 32 + lane;
            if (i < nr) {
                const uint32_t* __restrict__ rowp = reinterpret_cast<const uint32_t*>(px + i * S);
                uint32_t s1 = 0;
                unsigned long long s2 = 0;
                for (int t = 0; t < (S >> 1); t++) {   // pad samples are zero
                    const uint32_t w = rowp[t];
                    const uint32_t lo = w & 0xffffu, hi = w >> 16;
                    s1 += lo + hi;
                    s2 += (unsigned long long)lo * lo;
                    s2 += (unsigned long long)hi * hi;
                }
                // std along travel: sqrt(nc*S2 - S1^2) / (nc * D), exact integer numerator
                const double num = (double)((unsigned long long)nc * s2 - (unsigned long long)s1 * s1);
                sd[sl] = sqrt(num) / ((double)nc * Dd);
            }
        }
        double sd_max = warp_max(fmax(sd[0], sd[1]));
        double med_a, med_b;
        {
            int rank[2] = {0, 0};
            for (int t = 0; t < nr; t++) {
                const double o = __shfl_sync(0xffffffffu, (t >> 5) ? sd[1] : sd[0], t & 31);
#pragma unroll
                for (int sl = 0; sl < 2; sl++) {
                    const int me = sl * 32 + lane;
                    if (me < nr && (o < sd[sl] || (o == sd[sl] && t < me))) rank[sl]++;
                }
            }
            double ca = 0.0, cb = 0.0;
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int me = sl * 32 + lane;
                if (me < nr) {
                    if (rank[sl] == k1) ca = sd[sl];
                    if (rank[sl] == k2) cb = sd[sl];
                }
            }
            med_a = warp_sum(ca);   // exactly one lane holds each; the others contribute +0.0
            med_b = warp_sum(cb);
        }
        const double sd_med = (nr & 1) ? med_a : (med_a + med_b) / 2.0;
        const bool above = ((double)gmax / Dd) > c.p.height_threshold * f.picket_val[pk];
        const bool not_edge = sd_max < c.p.edge_threshold * sd_med;
        if (!(above && not_edge)) {
            if (lane == 0) { out.valid = 0; out.l = 0; out.r = 0; }
            continue;
        }
        // ---- 3. np.median(window, axis) -> 2 * median per travel sample (picketfence.py:1605-1609)
        uint32_t lmin = 0xffffffffu, lmax = 0;
        const int npairs = S >> 1;
        for (int t = lane; t < npairs; t += 32) {
            uint32_t m_lo, m_hi;
            if (nr <= 16) pair_median2<16>(px, S, nr, t, k1, k2, m_lo, m_hi);
            else if (nr <= 32) pair_median2<32>(px, S, nr, t, k1, k2, m_lo, m_hi);
            else pair_median2<64>(px, S, nr, t, k1, k2, m_lo, m_hi);
            if (2 * t < nc) { m2[2 * t] = m_lo; lmin = min(lmin, m_lo); lmax = max(lmax, m_lo); }
            if (2 * t + 1 < nc) { m2[2 * t + 1] = m_hi; lmin = min(lmin, m_hi); lmax = max(lmax, m_hi); }
        }
        lmin = warp_min(lmin);
        lmax = warp_max(lmax);
        __syncwarp();
        if (lmax == lmin) {  // flat profile: the reference divides by zero and then finds no peak
            if (lane == 0) { out.valid = 0; f.status = EPID_PF_WINDOW_NO_PEAK; }
            continue;
        }
        // ---- 4. FWXMProfilePhysical(ground=True, normalization=MAX) (core/profile.py:204-240); xs aliases px
        const double den = (double)(lmax - lmin);
        for (int j = lane; j < nc; j += 32) xs[j] = (double)(m2[j] - lmin) / den;
        __syncwarp();
        // find_peaks(values, fwxm_height=0.5, max_number=1) by prominence (core/profile.py:602-611, 2545-2623)
        double best_prom = -1.0;
        int best_idx = -1, best_lb = 0, best_rb = 0;
        for (int i = 1 + lane; i < nc - 1; i += 32) {
            if (xs[i - 1] < xs[i]) {
                int ahead = i + 1;
                while (ahead < nc - 1 && xs[ahead] == xs[i]) ahead++;
                if (xs[ahead] < xs[i]) {
                    const int p = (i + ahead - 1) / 2;
                    const double xp = xs[p];
                    int k = p, lb = p;
                    double lm = xp;
                    while (k >= 0 && xs[k] <= xp) { if (xs[k] < lm) { lm = xs[k]; lb = k; } k--; }
                    k = p;
                    int rb = p;
                    double rm = xp;
                    while (k <= nc - 1 && xs[k] <= xp) { if (xs[k] < rm) { rm = xs[k]; rb = k; } k++; }
                    const double prom = xp - fmax(lm, rm);
                    if (prom > best_prom || (prom == best_prom && p > best_idx)) { best_prom = prom; best_idx = p; best_lb = lb; best_rb = rb; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // warp arg-max by (prominence, index)
            const double op = __shfl_xor_sync(0xffffffffu, best_prom, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
            const int olb = __shfl_xor_sync(0xffffffffu, best_lb, o);
            const int orb = __shfl_xor_sync(0xffffffffu, best_rb, o);
            if (op > best_prom || (op == best_prom && oi > best_idx)) { best_prom = op; best_idx = oi; best_lb = olb; best_rb = orb; }
        }
        if (best_idx < 0) {
            if (lane == 0) { out.valid = 0; f.status = EPID_PF_WINDOW_NO_PEAK; }
            continue;
        }
        if (lane == 0) {
            const int p = best_idx;
            const double h = xs[p] - best_prom * 0.5;
            int k = p;
            while (best_lb < k && h < xs[k]) k--;
            double l = (double)k;
            if (xs[k] < h) l += (h - xs[k]) / (xs[k + 1] - xs[k]);
            k = p;
            while (k < best_rb && h < xs[k]) k++;
            double r = (double)k;
            if (xs[k] < h) r -= (h - xs[k]) / (xs[k - 1] - xs[k]);
            out.valid = 1;
            out.l = l;
            out.r = r;
        }
    }
}

int launch_pf_windows_fast(epid_ctx* ctx, cudaStream_t stream, const PfConst* cst, const FrameRef* refs, PfFrame* fr, PfWin* wins, int n) {
    static bool attr = false;
    const size_t smem = (sizeof(uint16_t) * W2_CAP + sizeof(uint32_t) * W2_MAXNC) * W2_WARPS;
    if (!attr) {
        EPID_CUDA(cudaFuncSetAttribute(k_pf_windows_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(W2_GRID_X, n);
    k_pf_windows_fast<<<grid, W2_WARPS * 32, smem, stream>>>(cst, refs, fr, wins);
    ctx->launches++;
    EPID_CUDA(cudaGetLastError());
    return EPID_OK;
}

}  // namespace epid
