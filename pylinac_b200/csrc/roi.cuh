// Point-in-quadrilateral rule shared by the ROI kernels (roi.cu, vmat.cu): the pixel selection of RectangleROI.pixels_flat
// (core/roi.py:641-660 -> skimage.draw.polygon): integer pixel coordinates inside the corner polygon or ON its boundary.
#pragma once

namespace epid {

// 0 outside, non-zero inside or on the boundary (crossing number with explicit edge / vertex tests, the structure of skimage's
// _geometry.point_in_polygon)
__device__ inline int point_in_quad(const double* vx, const double* vy, double x, double y) {
    int r = 0;
    double x0 = vx[3] - x, y0 = vy[3] - y;
    for (int i = 0; i < 4; i++) {
        const double x1 = vx[i] - x, y1 = vy[i] - y;
        if (y1 == 0 && (x1 == 0 || (y0 == 0 && ((x1 > 0) == (x0 < 0))))) return 2;      // vertex, or on a horizontal edge
        if ((y1 < 0) != (y0 < 0)) {      // the edge crosses the horizontal line through the point
            if (x0 >= 0) {
                if (x1 > 0) r += 1;                       // entirely to the right
                else {
                    const double det = (x0 * y1 - x1 * y0);
                    if (det == 0) return 3;               // on the edge
                    if ((det > 0) == (y1 > y0)) r += 1;
                }
            } else if (x1 > 0) {
                const double det = (x0 * y1 - x1 * y0);
                if (det == 0) return 3;
                if ((det > 0) == (y1 > y0)) r += 1;
            }
        }
        x0 = x1;
        y0 = y1;
    }
    return r & 1;
}

}  // namespace epid
