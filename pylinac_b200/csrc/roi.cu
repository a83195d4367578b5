// Region statistics on device-resident frames.
//
//   epid_roi_stats          RectangleROI.pixels_flat -> mean / std / min / max (core/roi.py:533-706): the pixels of a (possibly
//                           rotated) rectangle given by its four corners, selected like skimage.draw.polygon does -- integer pixel
//                           coordinates inside the polygon or ON its boundary (skimage's point_in_polygon returns non-zero for
//                           edge and vertex hits), clipped to the image.  skimage is not available in the build container: the
//                           rule is restated from its documentation / source as recalled (parity unpinned at that boundary); for
//                           axis-aligned rectangles with integer corners it reduces to a plain slice.
//   epid_weighted_centroid  WeightedCentroid.calculate (metrics/image.py:959-983): sum(idx * a) / sum(a) along both axes.
//
// One CTA per (frame, ROI).  Integer dtypes accumulate exact 64-bit sums (count, sum, sum of squares, index-weighted sums); float
// dtypes accumulate in fp64.  std = sqrt(mean(|x - mean|^2)) as numpy defines it, evaluated from the exact moments for integers.
#include <cmath>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "roi.cuh"

namespace epid {

constexpr int ROI_THREADS = 256;

template <typename T> struct RoiAcc { using type = double; };
template <> struct RoiAcc<uint8_t> { using type = unsigned long long; };
template <> struct RoiAcc<uint16_t> { using type = unsigned long long; };

struct RoiOut { double count, sum, sumsq, mn, mx, varnum; };   // varnum: exact N * S2 - S1^2 for 8 / 16-bit pixels, else -1

template <typename T>
__global__ void __launch_bounds__(ROI_THREADS)
k_roi_stats(const T* __restrict__ data, int H, int W, int nroi, const double* __restrict__ verts, RoiOut* __restrict__ out) {
    using A = typename RoiAcc<T>::type;
    const int fi = blockIdx.y, ri = blockIdx.x;
    const T* f = data + (size_t)fi * H * W;
    const double* v = verts + (size_t)ri * 8;          // (x, y) x 4
    const double vx[4] = {v[0], v[2], v[4], v[6]}, vy[4] = {v[1], v[3], v[5], v[7]};
    const double xmin = fmin(fmin(vx[0], vx[1]), fmin(vx[2], vx[3])), xmax = fmax(fmax(vx[0], vx[1]), fmax(vx[2], vx[3]));
    const double ymin = fmin(fmin(vy[0], vy[1]), fmin(vy[2], vy[3])), ymax = fmax(fmax(vy[0], vy[1]), fmax(vy[2], vy[3]));
    // skimage.draw._polygon: minr = int(max(0, r.min())), maxr = int(ceil(r.max())), clipped to shape - 1
    const int r0 = (int)fmax(0.0, ymin), r1 = min((int)ceil(ymax), H - 1);
    const int c0 = (int)fmax(0.0, xmin), c1 = min((int)ceil(xmax), W - 1);
    const int bh = r1 - r0 + 1, bw = c1 - c0 + 1;
    A s1 = 0, s2 = 0;
    unsigned long long cnt = 0;
    double mn = INFINITY, mx = -INFINITY;
    if (bh > 0 && bw > 0) {
        for (int i = threadIdx.x; i < bh * bw; i += ROI_THREADS) {
            const int r = r0 + i / bw, c = c0 + i % bw;
            if (!point_in_quad(vx, vy, (double)c, (double)r)) continue;
            const T pv = f[(size_t)r * W + c];
            const A a = (A)pv;
            s1 += a;
            s2 += a * a;
            cnt++;
            mn = fmin(mn, (double)pv);
            mx = fmax(mx, (double)pv);
        }
    }
    __shared__ A sh1[ROI_THREADS], sh2[ROI_THREADS];
    __shared__ unsigned long long shc[ROI_THREADS];
    __shared__ double shmn[ROI_THREADS], shmx[ROI_THREADS];
    sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2; shc[threadIdx.x] = cnt; shmn[threadIdx.x] = mn; shmx[threadIdx.x] = mx;
    __syncthreads();
    for (int s = ROI_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sh1[threadIdx.x] += sh1[threadIdx.x + s];
            sh2[threadIdx.x] += sh2[threadIdx.x + s];
            shc[threadIdx.x] += shc[threadIdx.x + s];
            shmn[threadIdx.x] = fmin(shmn[threadIdx.x], shmn[threadIdx.x + s]);
            shmx[threadIdx.x] = fmax(shmx[threadIdx.x], shmx[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        RoiOut o;
        o.count = (double)shc[0];
        o.sum = (double)sh1[0];
        o.sumsq = (double)sh2[0];
        o.mn = shmn[0];
        o.mx = shmx[0];
        o.varnum = -1.0;
        if (!std::is_floating_point<A>::value && shc[0] > 0) {
            // exact population variance numerator N * S2 - S1^2 in 128-bit integer arithmetic -> fp64 once
            const unsigned __int128 n = shc[0];
            const unsigned __int128 num = n * (unsigned __int128)(unsigned long long)sh2[0] -
                                          (unsigned __int128)(unsigned long long)sh1[0] * (unsigned long long)sh1[0];
            o.varnum = (double)num;
        }
        out[(size_t)fi * nroi + ri] = o;
    }
}

template <typename T>
static int do_roi(epid_ctx* ctx, const epid_batch* b, int nroi, const double* d_verts, RoiOut* d_out) {
    k_roi_stats<T><<<dim3(nroi, b->n), ROI_THREADS, 0, ctx->stream>>>((const T*)b->dptr, b->h, b->w, nroi, d_verts, d_out);
    ctx->launches++;
    EPID_CUDA(cudaGetLastError());
    return EPID_OK;
}

// ---------------------------------------------------------------------------------------- weighted centroid
template <typename T>
__global__ void __launch_bounds__(ROI_THREADS)
k_weighted_centroid(const T* __restrict__ data, int H, int W, double* __restrict__ part) {
    // grid (blocks, n): per-block partial sums of a, x * a, y * a (exact for integer pixels), combined on the host side of the C-ABI
    using A = typename RoiAcc<T>::type;
    const int fi = blockIdx.y;
    const T* f = data + (size_t)fi * H * W;
    A s = 0, sx = 0, sy = 0;
    const size_t per = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * ROI_THREADS + threadIdx.x; i < per; i += (size_t)gridDim.x * ROI_THREADS) {
        const A a = (A)f[i];
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        s += a;
        sx += a * (A)x;
        sy += a * (A)y;
    }
    __shared__ A sh[3][ROI_THREADS];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = sx; sh[2][threadIdx.x] = sy;
    __syncthreads();
    for (int k = ROI_THREADS / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) for (int j = 0; j < 3; j++) sh[j][threadIdx.x] += sh[j][threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        // integer sums travel as two 32-bit halves in doubles (exact), float sums as they are
        double* o = part + ((size_t)fi * gridDim.x + blockIdx.x) * 6;
        const A v = sh[threadIdx.x][0];
        if (std::is_floating_point<A>::value) { o[2 * threadIdx.x] = (double)v; o[2 * threadIdx.x + 1] = 0.0; }
        else {
            const unsigned long long u = (unsigned long long)v;
            o[2 * threadIdx.x] = (double)(u >> 32);
            o[2 * threadIdx.x + 1] = (double)(u & 0xffffffffull);
        }
    }
}

constexpr int WC_BLOCKS = 64;
template <typename T>
static int do_wc(epid_ctx* ctx, const epid_batch* b, double* d_part) {
    k_weighted_centroid<T><<<dim3(WC_BLOCKS, b->n), ROI_THREADS, 0, ctx->stream>>>((const T*)b->dptr, b->h, b->w, d_part);
    ctx->launches++;
    EPID_CUDA(cudaGetLastError());
    return EPID_OK;
}

}  // namespace epid

using namespace epid;

#define EPID_DISPATCH_ROI(dt, FN, ...)                                              \
    switch (dt) {                                                                   \
        case EPID_U8: rc = FN<uint8_t>(__VA_ARGS__); break;                         \
        case EPID_U16: rc = FN<uint16_t>(__VA_ARGS__); break;                       \
        case EPID_I16: rc = FN<int16_t>(__VA_ARGS__); break;                        \
        case EPID_I32: rc = FN<int32_t>(__VA_ARGS__); break;                        \
        case EPID_I64: rc = FN<long long>(__VA_ARGS__); break;                      \
        case EPID_F32: rc = FN<float>(__VA_ARGS__); break;                          \
        case EPID_F64: rc = FN<double>(__VA_ARGS__); break;                         \
        default: set_error("unknown dtype %d", dt); rc = EPID_ERR_INVALID;          \
    }

extern "C" int32_t epid_roi_stats(epid_ctx* ctx, const epid_batch* b, int32_t nroi, const double* verts_xy, double* count, double* mean,
                                  double* std, double* mn, double* mx) {
    EPID_REQUIRE(ctx && b && verts_xy && nroi > 0 && nroi <= 4096, EPID_ERR_INVALID, "bad argument");
    EPID_CUDA(cudaSetDevice(ctx->device));
    const size_t nv = sizeof(double) * 8 * nroi, no = sizeof(RoiOut) * (size_t)nroi * b->n;
    int rc = ensure_scratch(ctx, nv + no + 512);
    if (rc != EPID_OK) return rc;
    double* d_verts = (double*)ctx->scratch;
    RoiOut* d_out = (RoiOut*)((char*)ctx->scratch + (nv + 255) / 256 * 256);
    EPID_CUDA(cudaMemcpyAsync(d_verts, verts_xy, nv, cudaMemcpyHostToDevice, ctx->stream));
    EPID_DISPATCH_ROI(b->dtype, do_roi, ctx, b, nroi, d_verts, d_out);
    if (rc != EPID_OK) return rc;
    std::vector<RoiOut> h((size_t)nroi * b->n);
    EPID_CUDA(cudaMemcpyAsync(h.data(), d_out, no, cudaMemcpyDeviceToHost, ctx->stream));
    EPID_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < h.size(); i++) {
        const RoiOut& o = h[i];
        const double n = o.count;
        if (count) count[i] = n;
        const double m = n > 0 ? o.sum / n : NAN;
        if (mean) mean[i] = m;
        if (std) {
            if (!(n > 0)) std[i] = NAN;
            else if (o.varnum >= 0) std[i] = sqrt(o.varnum) / n;   // exact numerator
            else { const double var = o.sumsq / n - m * m; std[i] = var > 0 ? sqrt(var) : 0.0; }
        }
        if (mn) mn[i] = n > 0 ? o.mn : NAN;
        if (mx) mx[i] = n > 0 ? o.mx : NAN;
    }
    return EPID_OK;
}

extern "C" int32_t epid_weighted_centroid(epid_ctx* ctx, const epid_batch* b, double* cx, double* cy, double* total) {
    EPID_REQUIRE(ctx && b && cx && cy, EPID_ERR_INVALID, "NULL argument");
    EPID_CUDA(cudaSetDevice(ctx->device));
    const size_t np = sizeof(double) * 6 * WC_BLOCKS * (size_t)b->n;
    int rc = ensure_scratch(ctx, np + 256);
    if (rc != EPID_OK) return rc;
    double* d_part = (double*)ctx->scratch;
    EPID_DISPATCH_ROI(b->dtype, do_wc, ctx, b, d_part);
    if (rc != EPID_OK) return rc;
    std::vector<double> h((size_t)6 * WC_BLOCKS * b->n);
    EPID_CUDA(cudaMemcpyAsync(h.data(), d_part, np, cudaMemcpyDeviceToHost, ctx->stream));
    EPID_CUDA(cudaStreamSynchronize(ctx->stream));
    const bool integral = b->dtype == EPID_U8 || b->dtype == EPID_U16;
    for (int fi = 0; fi < b->n; fi++) {
        if (integral) {      // exact 128-bit totals of the 64 block partials, one fp64 division at the end like numpy's
            unsigned __int128 t[3] = {0, 0, 0};
            for (int k = 0; k < WC_BLOCKS; k++) {
                const double* o = &h[((size_t)fi * WC_BLOCKS + k) * 6];
                for (int j = 0; j < 3; j++) t[j] += ((unsigned __int128)(unsigned long long)o[2 * j] << 32) + (unsigned long long)o[2 * j + 1];
            }
            const double s = (double)t[0];
            if (total) total[fi] = s;
            cx[fi] = (double)t[1] / s;
            cy[fi] = (double)t[2] / s;
        } else {
            double t[3] = {0, 0, 0};
            for (int k = 0; k < WC_BLOCKS; k++) for (int j = 0; j < 3; j++) t[j] += h[((size_t)fi * WC_BLOCKS + k) * 6 + 2 * j];
            if (total) total[fi] = t[0];
            cx[fi] = t[1] / t[0];
            cy[fi] = t[2] / t[0];
        }
    }
    return EPID_OK;
}
