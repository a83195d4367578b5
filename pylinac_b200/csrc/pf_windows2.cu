// Two-kernel (leaf, picket) window path of the PicketFence pipeline: integer streaming work and the 1-D analysis are separated so
// that each runs with (nearly) all 32 lanes busy and without CTA barriers.
//
// Reference semantics (unchanged): PicketFence._get_mlc_window / _is_mlc_peak_in_window (picketfence.py:847-886) and
// MLCValue.get_peak_positions (picketfence.py:1605-1628) -> FWXMProfilePhysical.field_edge_idx (core/profile.py:602-611).
//
//   k_pf_win_medians   warp-autonomous.  A task = one leaf x a group of G neighbouring pickets (their windows are column ranges of
//       the same band of rows).  The band is copied RAW into a per-warp shared-memory slot by one cp.async.bulk (TMA, UBLKCP) per
//       row, completion on an mbarrier; two slots per warp, so the copy of the warp's next task is in flight while it works on the
//       current one.  P1: lanes own (row, picket) pairs -> sum / sum of squares / min / max of the row inside the picket's window;
//       the variance numerator nc * S2 - S1^2 is an exact integer and invariant under the frame's ground / inversion map, so raw
//       pixels do.  P2: lanes own pairs of band columns -> median over the rows by register sorting networks on packed u16x2
//       (VIMNMX.U16x2); the median commutes with the monotone ground / inversion map, which is applied to the result.  Output per
//       window (PfWinRec, HBM / L2): the median profile (2 * median in g units), the nr variance numerators and row extremes.
//   k_pf_win_fwxm      thread per window, 32 windows of a warp in lock step.  The numerators are ranked by a branch-free sorting
//       network (max and median of the row standard deviations -> _is_mlc_peak_in_window), then the serial integer FWXM analysis
//       of the median profile (lb_window_fwxm: fp64 only for the prominence, the half-height level and the two interpolations).
//       Records are transposed through shared memory ([sample][window], stride 33) so both the coalesced record reads and the
//       per-thread walks are bank-conflict free.
//
// Results are bit-identical to k_pf_windows_fast (same integer quantities, same fp64 expressions; tests/test_gpu_pf.py compares
// them).  Frames this path does not cover (Left-Right orientation, unaligned pitch, windows wider than 64 samples or taller than
// 32 rows, more than 1024 windows) are left to k_pf_windows_fast: this kernel sets PfFrame.win2 for the frames it takes.
#include <cstdio>
#include <cstdlib>

#include "pf_common.cuh"
#include "pf_win_common.cuh"
#include "tma.cuh"

namespace epid {

constexpr int WA_WARPS = 8;
constexpr int WA_SLOT = 6656;          // bytes per staging slot (26 rows x 256 B: two 51-sample windows of a 10 mm leaf at 2.56 px/mm)
constexpr int WA_GRID_X = 4;           // CTAs per frame
constexpr int WA_GMAX = 4;             // pickets per task
constexpr int WB_THREADS = 128;
constexpr int WB_ST = 33;              // transposed record stride (words): [sample][window of the warp]

__device__ __forceinline__ int wa_row_stride_bytes(int nvec) { return (nvec | 1) * 16; }   // odd vector count: rows start 4 banks apart

struct W2Geo {
    int ok, np, ninview, ntasks;
    int nvmax[WA_GMAX + 1];     // widest band (16-byte vectors per row) when pickets are taken g at a time
};

// LDGSTS = false: one cp.async.bulk (TMA) per band row, completion on an mbarrier.  LDGSTS = true: 16-byte cp.async per lane
// (LDGSTS), completion by cp.async.wait_group -- kept for comparison (EPID_WA_LOADER=1).
template <bool LDGSTS>
__global__ void __launch_bounds__(WA_WARPS * 32, 2)
k_pf_win_medians(const PfConst* __restrict__ cc, const FrameRef* __restrict__ frames, PfFrame* fr, PfWinRec* __restrict__ recs) {
    extern __shared__ __align__(128) unsigned char smraw[];          // WA_WARPS x 2 slots
    __shared__ __align__(8) unsigned long long s_bar[WA_WARPS][2];
    __shared__ W2Geo s_geo;
    __shared__ int s_a0[PF_P], s_a1[PF_P];
    __shared__ short s_gcs[WA_GMAX + 1][PF_P], s_gnv[WA_GMAX + 1][PF_P];   // per group size / group: first band column (view coordinates), vectors per row
    __shared__ short s_b0[PF_L], s_nr[PF_L];
    __shared__ unsigned char s_lg[PF_L];                              // pickets per task of this leaf
    __shared__ int s_toff[PF_L + 1];                                  // tasks before leaf li
    const int fi = blockIdx.y;
    const PfConst& c = *cc;
    PfFrame& f = fr[fi];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int H = c.H, W = c.W;
    const FrameRef frf = frames[fi];
    const int mis = (int)((reinterpret_cast<uintptr_t>(frf.origin) >> 1) & 7);
    if (lane == 0) {
        mbar_init(smem_u32(&s_bar[wid][0]), 1);
        mbar_init(smem_u32(&s_bar[wid][1]), 1);
    }
    if (tid == 0) mbar_fence_init();
    // ---- frame geometry: identical in every CTA of the frame (pure function of PfFrame / PfConst)
    const int st = f.status;
    const double sp = f.spacing;
    const int np = f.n_pickets, ninview = f.n_inview;
    const bool pre_ok = c.win2 && st == EPID_PF_OK && f.orientation == 0 && (frf.pitch & 7) == 0 && np >= 1 && np <= PF_P && ninview > 0 &&
                        ninview <= PF_L && sp == sp && sp >= 2.0 && sp < 4096.0 && (long long)ninview * np <= PF_W2_WCAP;
    if (!pre_ok) return;      // uniform across the CTA
    const double dpmm = c.p.dpmm;
    int bad = 0;
    for (int i = tid; i < ninview; i += WA_WARPS * 32) {
        const int leaf = f.inview[i];
        const double lw_px = c.p.leaf_width_mm[leaf] * dpmm;
        const double lc_px = c.p.leaf_center_mm[leaf] * dpmm + (double)H / 2.0;
        const int b0 = max((int)(lc_px - lw_px / 2.0), 0), b1 = min((int)(lc_px + lw_px / 2.0), H);   // _get_mlc_window: int() truncates
        s_b0[i] = (short)b0;
        s_nr[i] = (short)(b1 - b0);
        if (b1 - b0 > PF_W2_NRW) bad = 1;
    }
    if (wid == 0) {
        int a0 = W, a1 = 0;
        if (lane < np) {
            const double pidx = (double)f.picket_idx[lane];
            a0 = max((int)(pidx - sp / 2.0), 0);
            a1 = min((int)(pidx + sp / 2.0), W);
            s_a0[lane] = a0;
            s_a1[lane] = a1;
            if (a1 - a0 > PF_W2_NCW) bad = 1;
        }
        __syncwarp();
        for (int g = 1; g <= WA_GMAX; g++) {       // band of every group of g neighbouring pickets
            const int ng = (np + g - 1) / g;
            int nvec = 0, cs = 0;
            if (lane < ng) {
                int lo = W, hi = 0;
                for (int q = 0; q < g && lane * g + q < np; q++) {
                    const int x0 = s_a0[lane * g + q], x1 = s_a1[lane * g + q];
                    if (x1 > x0) { lo = min(lo, x0); hi = max(hi, x1); }
                }
                if (hi <= lo) { lo = 0; hi = 8; }
                cs = lo - ((lo + mis) & 7);
                const int ce = hi + ((8 - ((hi + mis) & 7)) & 7);
                nvec = (ce - cs) >> 3;
                s_gcs[g][lane] = (short)cs;
                s_gnv[g][lane] = (short)nvec;
            }
            const int nvmax = warp_max(nvec);
            if (lane == 0) s_geo.nvmax[g] = nvmax;
        }
    }
    bad = __syncthreads_or(bad);
    if (wid == 0) {
        // pickets per task of every leaf: as many as fit a slot with the leaf's row count; running task offsets
        int run = 0;
        for (int base = 0; base < ninview; base += 32) {
            const int i = base + lane;
            int cnt = 0;
            if (i < ninview) {
                const int nr = s_nr[i];
                int G = 0;
                for (int g = WA_GMAX; g >= 1 && G == 0; g--)
                    if (max(nr, 1) * wa_row_stride_bytes(s_geo.nvmax[g]) <= WA_SLOT) G = g;
                if (G == 0) bad = 1;
                s_lg[i] = (unsigned char)max(G, 1);
                cnt = (np + max(G, 1) - 1) / max(G, 1);
            }
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (i < ninview) s_toff[i] = run + inc - cnt;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            s_toff[ninview] = run;
            s_geo.ntasks = run;
            s_geo.ok = bad ? 0 : 1;
            if (!bad && blockIdx.x == 0) f.win2 = 1;
        }
    }
    __syncthreads();
    if (!s_geo.ok) return;
    const int ntasks = s_geo.ntasks;
    const int inv = f.inv;
    const uint32_t mn = f.mn, mx = f.mx;
    const int sag = c.p.sag_px;
    unsigned char* slot0 = smraw + (size_t)wid * 2 * WA_SLOT;
    const uint32_t bar0 = smem_u32(&s_bar[wid][0]), bar1 = smem_u32(&s_bar[wid][1]);
    PfWinRec* frecs = recs + (size_t)fi * PF_W2_WCAP;

    auto leaf_of = [&](int task, int li) { while (s_toff[li + 1] <= task) li++; return li; };     // tasks are visited in ascending order
    auto issue = [&](int task, int li, int sl) {
        const int G = s_lg[li], g = task - s_toff[li];
        const int b0 = s_b0[li], nr = s_nr[li];
        const int nvec = s_gnv[G][g], cs = s_gcs[G][g];
        if (nr <= 0) return;                         // nothing to copy: the consumer does not wait either
        const int RS = wa_row_stride_bytes(nvec);
        if (LDGSTS) {
            const int nv_tot = nr * nvec;
            const float inv_nvec = 1.0f / (float)nvec;
            for (int idx = lane; idx < nv_tot; idx += 32) {
                const int r = (int)(((float)idx + 0.5f) * inv_nvec), v = idx - r * nvec;
                int row = b0 + r - sag;
                if (sag) { row %= H; if (row < 0) row += H; }
                const uint32_t dst = smem_u32(slot0 + (size_t)sl * WA_SLOT + (size_t)r * RS + (size_t)v * 16);
                const void* src = frf.origin + ((ptrdiff_t)row * frf.pitch + cs + v * 8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
            return;
        }
        const uint32_t bar = sl ? bar1 : bar0;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)nr * (uint32_t)nvec * 16u);
        __syncwarp();
        if (lane < nr) {
            int row = b0 + lane - sag;               // np.roll(sag) folded into the source row
            if (sag) { row %= H; if (row < 0) row += H; }
            tma_load_1d(smem_u32(slot0 + (size_t)sl * WA_SLOT + (size_t)lane * RS), frf.origin + ((ptrdiff_t)row * frf.pitch + cs),
                        (uint32_t)nvec * 16u, bar);
        }
    };

    const int tstride = gridDim.x * WA_WARPS;
    int task = blockIdx.x * WA_WARPS + wid;
    uint32_t ph0 = 0, ph1 = 0;
    int cur = 0, li = 0, li_next = 0;
    if (task < ntasks) { li = leaf_of(task, 0); issue(task, li, 0); }
    if (LDGSTS) asm volatile("cp.async.commit_group;" ::: "memory");
    for (; task < ntasks; task += tstride, cur ^= 1, li = li_next) {
        if (task + tstride < ntasks) {
            li_next = leaf_of(task + tstride, li);
            issue(task + tstride, li_next, cur ^ 1);     // that slot was released by the __syncwarp at the end of the previous iteration
        }
        if (LDGSTS) asm volatile("cp.async.commit_group;" ::: "memory");      // (possibly empty) group of the next task
        const int G = s_lg[li], g = task - s_toff[li];
        const int nr = s_nr[li];
        const int nvec = s_gnv[G][g], cs = s_gcs[G][g];
        const int RS = wa_row_stride_bytes(nvec);
        const int Gn = min(G, np - g * G);
        PfWinRec* lrec = frecs + (size_t)li * np + (size_t)g * G;
        if (lane < Gn) {       // header: the shape of the window (empty windows are reported by the analysis kernel)
            const int a0 = s_a0[g * G + lane], a1 = s_a1[g * G + lane];
            lrec[lane].hdr = ((uint32_t)(uint16_t)(short)max(min(a1 - a0, 32767), -32768)) | ((uint32_t)(uint16_t)(short)nr << 16);
        }
        if (nr <= 0) { __syncwarp(); continue; }
        if (LDGSTS) { asm volatile("cp.async.wait_group 1;" ::: "memory"); __syncwarp(); }      // everything but the newest group has landed
        else if (cur) { mbar_wait(bar1, ph1); ph1 ^= 1u; } else { mbar_wait(bar0, ph0); ph0 ^= 1u; }
        const unsigned char* band = slot0 + (size_t)cur * WA_SLOT;
        // ---- P1: (row, picket) sums on the raw pixels
        const int ntk = nr * Gn;
        for (int tt = lane; tt < ntk; tt += 32) {
            int r, q;
            if (Gn == 1) { r = tt; q = 0; }
            else if (Gn == 2) { r = tt >> 1; q = tt & 1; }
            else if (Gn == 4) { r = tt >> 2; q = tt & 3; }
            else { r = tt / 3; q = tt - 3 * r; }
            const int a0 = s_a0[g * G + q], a1 = s_a1[g * G + q];
            unsigned long long numv = 0;
            uint32_t e = 0x0000ffffu;
            if (a1 > a0) {
                const unsigned char* rowp = band + (size_t)r * RS;
                int j0 = a0 - cs, j1 = a1 - cs;
                uint32_t s1 = 0, vmx = 0, vmn = 0xffffu;
                unsigned long long s2 = 0;
                if (j0 & 1) {
                    const uint32_t v = *reinterpret_cast<const uint16_t*>(rowp + 2 * j0);
                    s1 += v; s2 = mad_wide_u32(v, v, s2); vmx = max(vmx, v); vmn = min(vmn, v);
                    j0++;
                }
                if (j1 & 1) {
                    const uint32_t v = *reinterpret_cast<const uint16_t*>(rowp + 2 * (j1 - 1));
                    s1 += v; s2 = mad_wide_u32(v, v, s2); vmx = max(vmx, v); vmn = min(vmn, v);
                    j1--;
                }
                uint32_t mx2 = 0, mn2 = 0xffffffffu;
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(rowp);
                const int w0 = j0 >> 1, w1 = j1 >> 1;
                // rows r and r + 8 start in the same bank (the row stride is an odd number of 16-byte vectors): lanes start (r >> 3)
                // words into their window and wrap, so that no two lanes of the warp read the same bank
                const int rot = min(r >> 3, max(w1 - w0 - 1, 0));
                auto acc = [&](uint32_t x) {
                    const uint32_t lo = x & 0xffffu, hi = x >> 16;
                    s1 = __dp2a_lo(x, 0x0101u, s1);
                    s2 = mad_wide_u32(lo, lo, s2);
                    s2 = mad_wide_u32(hi, hi, s2);
                    mx2 = __vmaxu2(mx2, x);
                    mn2 = __vminu2(mn2, x);
                };
                const int cnt = w1 - w0;
#pragma unroll 4
                for (int t = 0; t < cnt; t++) {
                    int w = w0 + rot + t;
                    if (w >= w1) w -= cnt;
                    acc(wp[w]);
                }
                if (w1 > w0) {
                    vmx = max(vmx, max(mx2 & 0xffffu, mx2 >> 16));
                    vmn = min(vmn, min(mn2 & 0xffffu, mn2 >> 16));
                }
                const unsigned long long ncl = (unsigned long long)(a1 - a0);
                numv = ncl * s2 - (unsigned long long)s1 * s1;
                e = (vmx << 16) | vmn;
            }
            lrec[q].num[r] = numv;
            lrec[q].ext[r] = e;
        }
        // ---- P2: 2 * median over the rows for every pair of band columns between the group's first and last window
        {
            const uint16_t* px = reinterpret_cast<const uint16_t*>(band);
            const int S = RS >> 1;
            int lo = W, hi = 0;
            for (int q = 0; q < Gn; q++) {
                const int x0 = s_a0[g * G + q], x1 = s_a1[g * G + q];
                if (x1 > x0) { lo = min(lo, x0); hi = max(hi, x1); }
            }
            const int t_lo = (lo - cs) >> 1, t_hi = (hi - cs + 1) >> 1;
            for (int t = t_lo + lane; t < t_hi; t += 32) {
                const uint2 mm = pair_median_any(px, S, nr, t);
                const uint32_t g0 = inv ? 2u * mx - mm.x : mm.x - 2u * mn;
                const uint32_t g1 = inv ? 2u * mx - mm.y : mm.y - 2u * mn;
                const int c0 = 2 * t + cs;                 // view column of the low half
                for (int q = 0; q < Gn; q++) {
                    const int x0 = s_a0[g * G + q], x1 = s_a1[g * G + q];
                    if (c0 >= x0 && c0 < x1) lrec[q].m2[c0 - x0] = g0;
                    if (c0 + 1 >= x0 && c0 + 1 < x1) lrec[q].m2[c0 + 1 - x0] = g1;
                }
            }
        }
        __syncwarp();       // every lane is done with the slot: the next iteration may overwrite the other one... and this one after it
    }
}

// max, and the two middle order statistics of the nr keys a thread reads through key(i) (i < N slots, slots >= nr padded)
template <int N, class F>
__device__ __forceinline__ void rank_keys(F key, int nr, unsigned long long& kmax, unsigned long long& ka, unsigned long long& kb) {
    unsigned long long r[N];
    kmax = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        r[i] = i < nr ? key(i) : ~0ull;
        if (i < nr && r[i] > kmax) kmax = r[i];
    }
    sort_net_u64<N>(r);
    const int k1 = (nr - 1) / 2, k2 = nr / 2;
    ka = 0; kb = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (i == k1) ka = r[i];
        if (i == k2) kb = r[i];
    }
}

__global__ void __launch_bounds__(WB_THREADS)
k_pf_win_fwxm(const PfConst* __restrict__ cc, PfFrame* fr, const PfWinRec* __restrict__ recs, PfWin* __restrict__ wins) {
    __shared__ uint32_t s_buf[WB_THREADS / 32][PF_W2_NCW * WB_ST];
    const int fi = blockIdx.y;
    PfFrame& f = fr[fi];
    if (!f.win2) return;
    const PfConst& c = *cc;
    const int np = f.n_pickets;
    const int total = f.n_inview * np;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wbase = blockIdx.x * WB_THREADS + wid * 32;
    if (wbase >= total) return;
    const PfWinRec* frecs = recs + (size_t)fi * PF_W2_WCAP;
    uint32_t* buf = s_buf[wid];
    const int w = wbase + lane;
    const bool active = w < total;
    // ---- this thread's window: header, row extremes and variance numerators straight into registers (independent loads, all
    //      in flight at once; lanes read records 648 bytes apart)
    const PfWinRec& rec = frecs[active ? w : total - 1];
    const uint32_t hdr = rec.hdr;
    const int my_nc = (int)(short)(hdr & 0xffffu), my_nr = (int)(short)(hdr >> 16);
    int li = 0, pk = 0;
    if (active) { li = w / np; pk = w - li * np; }
    PfWin& out = wins[((size_t)fi * PF_L + li) * PF_P + pk];
    bool run = false;
    if (active) {
        if (my_nc <= 0 || my_nr <= 0) {           // empty slice: np.max raises ValueError in the reference
            out.valid = 0; out.l = 0; out.r = 0; f.status = EPID_PF_WINDOW_NO_PEAK;
        } else {
            run = true;
        }
    }
    const int nrr = run ? my_nr : 0;
    const int nr_all = __reduce_max_sync(0xffffffffu, nrr);
    // ---- _is_mlc_peak_in_window (picketfence.py:847-857): std along travel per row = sqrt(num) / (nc * D)
    unsigned long long kmax = 0, ka = 0, kb = 0;
    uint32_t my_vmx = 0, my_vmn = 0xffffu;
    {
        auto key = [&](int i) { return rec.num[i]; };
        if (nr_all <= 16) rank_keys<16>(key, nrr, kmax, ka, kb);       // warp-uniform choice
        else rank_keys<32>(key, nrr, kmax, ka, kb);
#pragma unroll
        for (int i = 0; i < PF_W2_NRW; i++) {
            if (i < nrr) {
                const uint32_t e = rec.ext[i];
                my_vmx = max(my_vmx, e >> 16);
                my_vmn = min(my_vmn, e & 0xffffu);
            }
        }
    }
    if (run) {
        const double Dd = (double)f.D;
        const double dn = (double)my_nc * Dd;
        const double sd_max = sqrt((double)kmax) / dn;
        const double sa = sqrt((double)ka) / dn, sb = sqrt((double)kb) / dn;
        const double sd_med = (my_nr & 1) ? sa : (sa + sb) / 2.0;
        const uint32_t gmax = f.inv ? f.mx - my_vmn : my_vmx - f.mn;
        const bool above = ((double)gmax / Dd) > c.p.height_threshold * f.picket_val[pk];
        const bool not_edge = sd_max < c.p.edge_threshold * sd_med;
        if (!(above && not_edge)) {
            out.valid = 0; out.l = 0; out.r = 0;
            run = false;
        }
    }
    __syncwarp();
    // ---- median profiles of the 32 windows -> [sample][window]
    {
        const int nc_all = __reduce_max_sync(0xffffffffu, run ? my_nc : 0);     // samples beyond a window's own nc are never read
#pragma unroll 8
        for (int k = 0; k < 32; k++) {
            const PfWinRec& rc = frecs[min(wbase + k, total - 1)];
            const uint32_t v0 = rc.m2[lane];
            const uint32_t v1 = nc_all > 32 ? rc.m2[lane + 32] : 0u;
            buf[lane * WB_ST + k] = v0;
            if (nc_all > 32) buf[(lane + 32) * WB_ST + k] = v1;
        }
    }
    __syncwarp();
    if (run) {
        double l = 0, r = 0;
        const int v = lb_window_fwxm<WB_ST>(buf + lane, my_nc, l, r);
        out.valid = v;
        if (v) { out.l = l; out.r = r; }
        else f.status = EPID_PF_WINDOW_NO_PEAK;
    }
}

size_t pf_win2_scratch_bytes(int n) { return sizeof(PfWinRec) * (size_t)n * PF_W2_WCAP; }

int launch_pf_windows2(epid_ctx* ctx, cudaStream_t stream, const PfConst* cst, const FrameRef* refs, PfFrame* fr, PfWinRec* recs, PfWin* wins,
                       int n, PfTimers* tm) {
    const size_t smem = (size_t)WA_WARPS * 2 * WA_SLOT;
    EPID_SMEM_OPT_IN(ctx, k_pf_win_medians<false>, smem);
    static int gx = 0;
    if (gx == 0) { const char* e = getenv("EPID_WA_GRID"); gx = e ? atoi(e) : WA_GRID_X; if (gx < 1 || gx > 64) gx = WA_GRID_X; }
    static int loader = -1;
    if (loader < 0) { const char* e = getenv("EPID_WA_LOADER"); loader = e ? atoi(e) : 0; }
    if (loader == 1) {
        EPID_SMEM_OPT_IN(ctx, k_pf_win_medians<true>, smem);
        k_pf_win_medians<true><<<dim3(gx, n), WA_WARPS * 32, smem, stream>>>(cst, refs, fr, recs);
    } else {
        k_pf_win_medians<false><<<dim3(gx, n), WA_WARPS * 32, smem, stream>>>(cst, refs, fr, recs);
    }
    ctx->launches++;
    if (tm) { int rc = tm->mark(stream, PF_STAGE_WIN_MEDIANS); if (rc != EPID_OK) return rc; }
    k_pf_win_fwxm<<<dim3(PF_W2_WCAP / WB_THREADS, n), WB_THREADS, 0, stream>>>(cst, fr, recs, wins);
    ctx->launches++;
    if (tm) { int rc = tm->mark(stream, PF_STAGE_WIN_FWXM); if (rc != EPID_OK) return rc; }
    EPID_CUDA(cudaGetLastError());
    return EPID_OK;
}

}  // namespace epid
