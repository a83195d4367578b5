// Last stage of the batched PicketFence pipeline: one CTA per frame turns the per-window FWHM edges into the
// measurement table and the PFResult scalars.
//
// Reference semantics: leaf-row pruning by the median kiss count (picketfence.py:810-828), MLCValue.get_peak_positions /
// error / marker_lines (picketfence.py:1605-1628, 1701-1743), Picket.get_fit / dist2cax / skew (picketfence.py:1881-1923),
// aggregates (picketfence.py:439-562, 1313-1363, 1467-1469).
//
// Everything that scales with the number of measurements runs data-parallel: thread per (leaf, picket) pair for the table
// and the errors (the table index of a pair is its leaf row's offset plus a popcount over the row's validity mask), warp per
// picket for the line fits and the width statistics, block reductions for the aggregates.  Only O(leaves) and O(pickets)
// bookkeeping is left to one thread.
#include "pf_common.cuh"

namespace epid {

constexpr int FIN_WARPS = FIN_THREADS / 32;
constexpr int FIN_LPL = (PF_L + 31) / 32;       // leaf slots per lane

// np.median of n non-negative doubles a[0..n) (shared memory) by the whole block: radix selection of the lower middle order
// statistic on the bit patterns (for non-negative doubles the unsigned order of the bits is the order of the values), eight 8-bit
// digits, warp-aggregated shared-memory histograms; the upper middle value is the same value if it occurs often enough, else the
// smallest larger one.  Thread 0 returns the median; every thread must call.
__device__ inline double block_median_nonneg_f64(const double* __restrict__ a, int n) {
    __shared__ uint32_t s_hist[256];
    __shared__ unsigned long long s_prefix, s_above;
    __shared__ uint32_t s_k, s_cle;
    const int tid = threadIdx.x, lane = tid & 31;
    auto keyat = [&](int i) { return (unsigned long long)__double_as_longlong(a[i]); };
    const int nround = (n + FIN_THREADS - 1) / FIN_THREADS * FIN_THREADS;      // every thread takes the same number of trips
    const int k1 = (n - 1) / 2, k2 = n / 2;
    unsigned long long prefix = 0, mask = 0;
    uint32_t k = (uint32_t)k1;
    for (int shift = 56; shift >= 0; shift -= 8) {
        s_hist[tid & 255] = 0;
        __syncthreads();
        for (int i = tid; i < nround; i += FIN_THREADS) {
            const unsigned long long kv = i < n ? keyat(i) : 0ull;
            const bool on = i < n && (kv & mask) == prefix;
            const uint32_t d = (uint32_t)(kv >> shift) & 255u;
            const unsigned m = __match_any_sync(0xffffffffu, on ? d : 256u);
            if (on && lane == __ffs(m) - 1) atomicAdd(&s_hist[d], (uint32_t)__popc(m));
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) { loc[e] = s_hist[tid * 8 + e]; sum += loc[e]; }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            uint32_t cum = inc - sum;
            if (k >= cum && k < inc) {      // exactly one lane
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    if (k >= cum && k < cum + loc[e]) { s_k = k - cum; s_prefix = prefix | ((unsigned long long)(tid * 8 + e) << shift); }
                    cum += loc[e];
                }
            }
        }
        __syncthreads();
        prefix = s_prefix;
        k = s_k;
        mask |= 0xffull << shift;
    }
    // upper middle value
    if (tid == 0) { s_cle = 0; s_above = ~0ull; }
    __syncthreads();
    uint32_t c = 0;
    unsigned long long above = ~0ull;
    for (int i = tid; i < n; i += FIN_THREADS) {
        const unsigned long long kv = keyat(i);
        c += kv <= prefix ? 1u : 0u;
        if (kv > prefix && kv < above) above = kv;
    }
    c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, above, o); above = t < above ? t : above; }
    if (lane == 0) { atomicAdd(&s_cle, c); atomicMin(&s_above, above); }
    __syncthreads();
    const double v1 = __longlong_as_double((long long)prefix);
    const double v2 = (s_cle >= (uint32_t)k2 + 1u) ? v1 : __longlong_as_double((long long)s_above);
    return (n & 1) ? v1 : (v1 + v2) / 2.0;
}

__global__ void __launch_bounds__(FIN_THREADS)
k_pf_finalize(const PfConst* __restrict__ cc, PfFrame* fr, const PfWin* __restrict__ wins, epid_pf_summary* __restrict__ summ,
              epid_pf_meas* __restrict__ meas_all) {
    extern __shared__ double s_err[];                    // pow2(2 * meas_cap) doubles for the median of |errors|
    __shared__ int s_cnt[PF_L], s_off[PF_L], s_keep[PF_L];
    __shared__ uint32_t s_vmask[PF_L];
    __shared__ double s_upper[PF_L], s_centre[PF_L];
    __shared__ double s_fit[PF_P][2];
    __shared__ int s_i[8];
    __shared__ double s_wbuf[FIN_WARPS][PF_L], s_wsort[FIN_WARPS][PF_L];
    __shared__ double s_rmax[FIN_WARPS];
    __shared__ int s_rarg[FIN_WARPS], s_rpass[FIN_WARPS], s_rfail[FIN_WARPS];

    const int fi = blockIdx.x;
    const PfConst& c = *cc;
    PfFrame& f = fr[fi];
    epid_pf_summary& S = summ[fi];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int H = c.H, W = c.W;
    // every word of the row is defined, whatever the frame's fate: rows of different runs / pipelines compare equal byte for byte
    for (int k = tid; k < (int)(sizeof(epid_pf_summary) / 4); k += FIN_THREADS) reinterpret_cast<uint32_t*>(&S)[k] = 0u;
    __syncthreads();
    if (tid == 0) {
        S.status = f.status;
        S.orientation = f.orientation;
        S.noise_median_passes = f.noise_passes;
        S.corner_inverted = f.corner_inverted;
        S.height = H;
        S.width = W;
        S.n_pickets = f.n_pickets;
        S.n_meas = 0;
        S.n_leaves_removed = 0;
        S.picket_spacing_px = f.spacing;
    }
    if (tid < PF_P) {
        S.picket_idx[tid] = tid < f.n_pickets ? f.picket_idx[tid] : 0;
        S.picket_val[tid] = tid < f.n_pickets ? f.picket_val[tid] : 0.0;
    }
    if (f.status != EPID_PF_OK) return;
    const int nl = f.n_inview, np = f.n_pickets;
    const int orient = f.orientation;
    const int npos = c.p.separate_leaves ? 2 : 1;
    const double dpmm = c.p.dpmm;
    const double spacing = f.spacing;
    const PfWin* wf = wins + (size_t)fi * PF_L * PF_P;
    const double n_axis_half = (orient == 0 ? (double)H : (double)W) / 2.0;
    const double ratio = c.p.leaf_analysis_width_ratio;
    // ---- kisses per leaf row, marker-line geometry of the row (picketfence.py:1725-1743)
    for (int l = tid; l < nl; l += FIN_THREADS) {
        uint32_t m = 0;
        for (int p = 0; p < np; p++) m |= (wf[l * PF_P + p].valid ? 1u : 0u) << p;
        s_vmask[l] = m;
        s_cnt[l] = __popc(m);
        const int leaf = f.inview[l];
        const double lw_px = c.p.leaf_width_mm[leaf] * dpmm;
        const double lc_px = c.p.leaf_center_mm[leaf] * dpmm + n_axis_half;
        const double upper = lc_px - lw_px / 2.0 * ratio;
        const double lower = lc_px + lw_px / 2.0 * ratio;
        s_upper[l] = upper;
        s_centre[l] = (lower - upper) / 2.0 + upper;          // Line.center (core/geometry.py:556-561)
    }
    __syncthreads();
    if (tid == 0) {
        // median over the leaf rows that have at least one measurement (group_by on mlc_meas, picketfence.py:810-814);
        // counts are <= 32, so a counting sort gives the two middle order statistics
        int hist[PF_P + 1];
        for (int k = 0; k <= PF_P; k++) hist[k] = 0;
        int ng = 0, total = 0;
        for (int l = 0; l < nl; l++)
            if (s_cnt[l] > 0) { hist[s_cnt[l]]++; ng++; total += s_cnt[l]; }
        int status = EPID_PF_OK;
        int removed = 0;
        if (total == 0) {
            status = EPID_PF_NO_MEASUREMENTS;
        } else {
            const int ka = (ng & 1) ? ng / 2 : ng / 2 - 1, kb = ng / 2;
            int va = 0, vb = 0, acc = 0;
            bool ha = false, hb = false;
            for (int k = 1; k <= PF_P; k++) {
                acc += hist[k];
                if (!ha && acc > ka) { va = k; ha = true; }
                if (!hb && acc > kb) { vb = k; hb = true; }
            }
            const int med_twice = va + vb;                     // 2 * statistics.median
            int off = 0;
            for (int l = 0; l < nl; l++) {
                const bool keep = s_cnt[l] > 0 && 2 * s_cnt[l] == med_twice;
                s_keep[l] = keep ? 1 : 0;
                s_off[l] = off;
                if (keep) off += s_cnt[l];
                else if (s_cnt[l] > 0) removed++;
            }
            if (off == 0) status = EPID_PF_NO_MEASUREMENTS;     // a .5 median drops every row (reference: polyfit of nothing)
            else if (off > c.meas_cap) status = EPID_PF_CAPACITY;
            s_i[1] = off;
        }
        s_i[0] = status;
        S.n_leaves_removed = removed;
        if (status != EPID_PF_OK) { S.status = status; f.status = status; }
    }
    __syncthreads();
    if (s_i[0] != EPID_PF_OK) return;
    const int M = s_i[1];
    epid_pf_meas* meas = meas_all + (size_t)fi * c.meas_cap;
    // ---- measurement table, leaf-major / picket-minor (= PicketFence.mlc_meas order): thread per (leaf, picket)
    const int npairs = nl * np;
    for (int t = tid; t < npairs; t += FIN_THREADS) {
        const int l = t / np, p = t - l * np;
        const uint32_t vm = s_vmask[l];
        if (!s_keep[l] || !((vm >> p) & 1u)) continue;
        const int q = s_off[l] + __popc(vm & ((1u << p) - 1u));
        const PfWin w = wf[l * PF_P + p];
        epid_pf_meas& m = meas[q];
        m.leaf_num = c.p.leaf_num[f.inview[l]];
        m.picket = p;
        const double offp = fmax((double)f.picket_idx[p] - spacing / 2.0, 0.0);   // picketfence.py:1618-1627
        if (npos == 2) {
            m.position[0] = w.l + offp;
            m.position[1] = w.r + offp;
        } else {
            m.position[0] = fabs(w.r - w.l) / 2.0 + w.l + offp;                       // center_idx (core/profile.py:322-327)
            m.position[1] = 0.0;
        }
        m.width_mm = (fmax(w.r, w.l) - fmin(w.r, w.l)) / dpmm;                        // field_width_px / dpmm
    }
    // ---- per-picket line fit np.polyfit(along-leaf-stack, along-travel, 1)  (picketfence.py:1881-1899): warp per picket
    for (int p = wid; p < np; p += FIN_WARPS) {
        const double offp = fmax((double)f.picket_idx[p] - spacing / 2.0, 0.0);
        double xv[FIN_LPL], y0[FIN_LPL], y1[FIN_LPL];
        bool on[FIN_LPL];
        double sx = 0, sy = 0;
        int n = 0;
#pragma unroll
        for (int it = 0; it < FIN_LPL; it++) {
            const int l = it * 32 + lane;
            on[it] = l < nl && s_keep[l] && ((s_vmask[l] >> p) & 1u);
            xv[it] = 0; y0[it] = 0; y1[it] = 0;
            if (on[it]) {
                const PfWin w = wf[l * PF_P + p];
                xv[it] = s_upper[l];
                if (npos == 2) {
                    y0[it] = w.l + offp; y1[it] = w.r + offp;
                    sx += xv[it] * 2.0; sy += y0[it] + y1[it]; n += 2;
                } else {
                    y0[it] = fabs(w.r - w.l) / 2.0 + w.l + offp;
                    sx += xv[it]; sy += y0[it]; n += 1;
                }
            }
        }
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        n = warp_sum(n);
        if (n == 0) {
            if (lane == 0) { s_fit[p][0] = __longlong_as_double(0x7ff8000000000000LL); s_fit[p][1] = s_fit[p][0]; }
            continue;
        }
        const double mx_ = sx / n, my_ = sy / n;
        double sxx = 0, sxy = 0;
#pragma unroll
        for (int it = 0; it < FIN_LPL; it++) {
            if (on[it]) {
                const double dx = xv[it] - mx_;
                if (npos == 2) { sxx += 2.0 * dx * dx; sxy += dx * (y0[it] - my_) + dx * (y1[it] - my_); }
                else { sxx += dx * dx; sxy += dx * (y0[it] - my_); }
            }
        }
        sxx = warp_sum(sxx);
        sxy = warp_sum(sxy);
        if (lane == 0) {
            const double slope = sxx > 0 ? sxy / sxx : 0.0;
            s_fit[p][0] = slope;
            s_fit[p][1] = my_ - slope * mx_;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int bad = 0;
        for (int p = 0; p < np; p++)
            if (s_fit[p][0] != s_fit[p][0]) bad = 1;   // a picket without measurements: polyfit([]) raises
        s_i[2] = bad;
        if (bad) { S.status = 7; f.status = 7; }
    }
    __syncthreads();
    if (s_i[2]) return;
    // ---- errors (picketfence.py:1701-1718) + per-thread partial aggregates
    int m2n = 1;
    while (m2n < M * npos) m2n <<= 1;
    for (int q = M * npos + tid; q < m2n; q += FIN_THREADS) s_err[q] = __longlong_as_double(0x7ff0000000000000LL);
    int t_pass = 0, t_failed = 0, t_arg = 0x7fffffff;
    double t_max = -1.0;
    for (int t = tid; t < npairs; t += FIN_THREADS) {
        const int l = t / np, p = t - l * np;
        const uint32_t vm = s_vmask[l];
        if (!s_keep[l] || !((vm >> p) & 1u)) continue;
        const int q = s_off[l] + __popc(vm & ((1u << p) - 1u));
        epid_pf_meas& m = meas[q];
        const double fitv = s_fit[p][0] * s_centre[l] + s_fit[p][1];
        double me = 0.0;
        bool allp = true;
        for (int s = 0; s < npos; s++) {
            double picket_pos = fitv;
            if (npos == 2) picket_pos += (s == 0 ? -1.0 : 1.0) * c.p.nominal_gap_mm / 2.0 * dpmm;
            const double e = (m.position[s] - picket_pos) / dpmm;
            const int ok = fabs(e) < c.p.tolerance ? 1 : 0;
            m.error[s] = e;
            m.passed[s] = ok;
            s_err[q * npos + s] = fabs(e);
            t_pass += ok;
            if (!ok) allp = false;
            me = fmax(me, fabs(e));
        }
        if (npos == 1) { m.error[1] = 0.0; m.passed[1] = 1; }
        if (!allp) t_failed++;
        if (me > t_max) { t_max = me; t_arg = q; }     // q grows with t: first maximum of this thread's subsequence
    }
    // block reduction; first maximum in table order = stable descending sort .first()
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(0xffffffffu, t_max, o);
        const int oa = __shfl_xor_sync(0xffffffffu, t_arg, o);
        if (om > t_max || (om == t_max && oa < t_arg)) { t_max = om; t_arg = oa; }
    }
    t_pass = warp_sum(t_pass);
    t_failed = warp_sum(t_failed);
    if (lane == 0) { s_rmax[wid] = t_max; s_rarg[wid] = t_arg; s_rpass[wid] = t_pass; s_rfail[wid] = t_failed; }
    __syncthreads();
    // ---- aggregates
    if (tid == 0) {
        int n_pass = 0, n_failed = 0, arg = s_rarg[0];
        double max_err = s_rmax[0];
        for (int k = 0; k < FIN_WARPS; k++) {
            n_pass += s_rpass[k];
            n_failed += s_rfail[k];
            if (s_rmax[k] > max_err || (s_rmax[k] == max_err && s_rarg[k] < arg)) { max_err = s_rmax[k]; arg = s_rarg[k]; }
        }
        const int n_tot = M * npos;
        S.n_meas = M;
        S.percent_passing = 100.0 * (double)n_pass / (double)n_tot;
        S.max_error_mm = max_err;
        S.max_error_picket = meas[arg].picket;
        S.max_error_leaf = meas[arg].leaf_num;
        S.max_error_bank = (npos == 2 && !(fabs(meas[arg].error[0]) > fabs(meas[arg].error[1]))) ? 1 : 0;
        S.passed = n_pass == n_tot ? 1 : 0;
        S.n_failed = n_failed;
        // dist2cax (picketfence.py:1905-1923) / image.center (core/image.py:526-533, PFDicomImage.center :246-260)
        double cax;
        if (c.p.has_cax_override) cax = orient == 0 ? c.p.cax_x_px : c.p.cax_y_px;
        else cax = (orient == 0 ? (double)W : (double)H) / 2.0 - 0.5;
        S.cax_px = cax;
        const int length = orient == 0 ? H : W;
        const double xmid = rint((double)length / 2.0);
        double d2c[PF_P], srt[PF_P];
        double skew = 0.0;
        for (int p = 0; p < np; p++) {
            S.fit_slope[p] = s_fit[p][0];
            S.fit_intercept[p] = s_fit[p][1];
            d2c[p] = (cax - (s_fit[p][0] * xmid + s_fit[p][1])) / dpmm;
            S.offsets_from_cax_mm[p] = d2c[p];
            skew += s_fit[p][0] * (180.0 / 3.14159265358979323846);
            int j = p;
            while (j > 0 && srt[j - 1] > d2c[p]) { srt[j] = srt[j - 1]; j--; }
            srt[j] = d2c[p];
        }
        S.mlc_skew = skew / (double)np;
        double sp = 0.0;
        for (int p = 0; p + 1 < np; p++) sp += fabs(srt[p] - srt[p + 1]);
        S.mean_picket_spacing_mm = np > 1 ? sp / (double)(np - 1) : __longlong_as_double(0x7ff8000000000000LL);
    }
    // ---- picket widths (picketfence.py:471-491): warp per picket, rank sort of <= 160 widths in shared memory
    for (int p = wid; p < np; p += FIN_WARPS) {
        double* wb = s_wbuf[wid];
        double* ws = s_wsort[wid];
        int n = 0;
#pragma unroll
        for (int it = 0; it < FIN_LPL; it++) {
            const int l = it * 32 + lane;
            const bool on = l < nl && s_keep[l] && ((s_vmask[l] >> p) & 1u);
            const unsigned b = __ballot_sync(0xffffffffu, on);
            if (on) {
                const PfWin w = wf[l * PF_P + p];
                wb[n + __popc(b & ((1u << lane) - 1u))] = (fmax(w.r, w.l) - fmin(w.r, w.l)) / dpmm;
            }
            n += __popc(b);
        }
        __syncwarp();
        for (int e = lane; e < n; e += 32) {
            const double v = wb[e];
            int rank = 0;
            for (int j = 0; j < n; j++) {
                const double o = wb[j];
                rank += (o < v || (o == v && j < e)) ? 1 : 0;
            }
            ws[rank] = v;
        }
        __syncwarp();
        if (lane == 0) {
            double sum = 0.0;
            for (int j = 0; j < n; j++) sum += wb[j];      // table order, like the reference's list
            S.picket_width_max[p] = ws[n - 1];
            S.picket_width_min[p] = ws[0];
            S.picket_width_mean[p] = sum / (double)n;
            S.picket_width_median[p] = (n & 1) ? ws[n / 2] : (ws[n / 2 - 1] + ws[n / 2]) / 2.0;
        }
        __syncwarp();
    }
    // ---- median of |errors| (np.median)
    __syncthreads();
    {
        const int ne = M * npos;
        const double med = block_median_nonneg_f64(s_err, ne);
        if (tid == 0) S.abs_median_error_mm = med;
    }
}

int launch_pf_finalize(epid_ctx* ctx, cudaStream_t stream, const PfConst* cst, PfFrame* fr, const PfWin* wins, epid_pf_summary* summ,
                       epid_pf_meas* meas, int n, int meas_cap) {
    int m2 = 1;
    while (m2 < 2 * meas_cap) m2 <<= 1;
    const size_t smem = sizeof(double) * m2;
    if (smem > 16 * 1024) EPID_SMEM_OPT_IN(ctx, k_pf_finalize, smem);
    k_pf_finalize<<<n, FIN_THREADS, smem, stream>>>(cst, fr, wins, summ, meas);
    ctx->launches++;
    EPID_CUDA(cudaGetLastError());
    return EPID_OK;
}

}  // namespace epid
