// BaseImage.gamma (core/image.py:928-1017): Bakai gamma map between a reference and a comparison image.
//
//   ref[ref < threshold * max(ref)] = nan
//   img_x = ndimage.sobel(float32(ref), 1); img_y = ndimage.sobel(float32(ref), 0); grad = np.hypot(img_x, img_y)      (float32)
//   gamma = |comp - ref| / sqrt((doseTA / 100)^2 + distTA_pixels^2 * grad^2)                                           (float64 / float32)
//
// One fused kernel: every output pixel evaluates its 3 x 3 Sobel stencil from the float64 reference directly.  scipy's sobel is two
// correlate1d passes (derivative [-1, 0, 1], then smoothing [1, 2, 1], mode 'reflect'), each accumulating in double and storing
// float32: the per-pass rounding is reproduced (the derivative values are rounded to float32 before they are smoothed).  The zero
// centre weight of the derivative still multiplies the centre sample (0 * nan = nan), as in ni_filters.c.  The denominator follows
// numpy's float32 arithmetic (python scalars are weak): sqrtf(f32(dose^2) + f32(dist^2) * (grad * grad)).
#include <cmath>

#include "common.cuh"

namespace epid {

__device__ __forceinline__ int reflect_idx(int i, int n) {      // scipy 'reflect': d c b a | a b c d | d c b a
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i;
}

__global__ void k_gamma(const double* __restrict__ ref, const double* __restrict__ comp, int H, int W, double thr_abs, float dose2, float dist2,
                        double* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, fi = blockIdx.z;
    if (x >= W || y >= H) return;
    const double* r = ref + (size_t)fi * H * W;
    const double* c = comp + (size_t)fi * H * W;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    auto in32 = [&](int yy, int xx) -> float {
        const double v = r[(size_t)reflect_idx(yy, H) * W + reflect_idx(xx, W)];
        return (float)(v < thr_abs ? qnan : v);
    };
    float dx[3], dy[3];
#pragma unroll
    for (int j = -1; j <= 1; j++) {
        // derivative along x at rows y + j (first pass of sobel(axis=1)), along y at columns x + j (first pass of sobel(axis=0))
        const double cx = (double)in32(y + j, x) * 0.0, cy = (double)in32(y, x + j) * 0.0;
        dx[j + 1] = (float)(cx + ((double)in32(y + j, x - 1) - (double)in32(y + j, x + 1)) * -1.0);
        dy[j + 1] = (float)(cy + ((double)in32(y - 1, x + j) - (double)in32(y + 1, x + j)) * -1.0);
    }
    const float img_x = (float)((double)dx[1] * 2.0 + ((double)dx[0] + (double)dx[2]) * 1.0);
    const float img_y = (float)((double)dy[1] * 2.0 + ((double)dy[0] + (double)dy[2]) * 1.0);
    const float grad = hypotf(img_x, img_y);
    const float den = sqrtf(dose2 + dist2 * (grad * grad));
    const double rv = r[(size_t)y * W + x];
    const double rn = rv < thr_abs ? qnan : rv;
    out[(size_t)fi * H * W + (size_t)y * W + x] = fabs(c[(size_t)y * W + x] - rn) / (double)den;
}

}  // namespace epid

using namespace epid;

extern "C" int32_t epid_gamma(epid_ctx* ctx, const epid_batch* ref, const epid_batch* comp, double threshold_abs, double dose_frac,
                              double dist_px, epid_batch** out) {
    EPID_REQUIRE(ctx && ref && comp && out, EPID_ERR_INVALID, "NULL argument");
    EPID_REQUIRE(ref->dtype == EPID_F64 && comp->dtype == EPID_F64, EPID_ERR_UNSUPPORTED, "gamma takes float64 (ground / normalised) images");
    EPID_REQUIRE(ref->n == comp->n && ref->h == comp->h && ref->w == comp->w, EPID_ERR_INVALID, "The images are not the same size");
    EPID_CUDA(cudaSetDevice(ctx->device));
    int rc = epid_batch_alloc(ctx, EPID_F64, ref->n, ref->h, ref->w, out);
    if (rc != EPID_OK) return rc;
    const dim3 block(32, 8), grid((ref->w + 31) / 32, (ref->h + 7) / 8, ref->n);
    // numpy: python-float scalars are weak next to a float32 array -> both constants are rounded to float32 once
    const float dose2 = (float)(dose_frac * dose_frac), dist2 = (float)(dist_px * dist_px);
    k_gamma<<<grid, block, 0, ctx->stream>>>((const double*)ref->dptr, (const double*)comp->dptr, ref->h, ref->w, threshold_abs, dose2, dist2,
                                             (double*)(*out)->dptr);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { set_error("gamma kernel failed: %s", cudaGetErrorString(e)); epid_batch_free(*out); *out = nullptr; return EPID_ERR_CUDA; }
    return EPID_OK;
}
