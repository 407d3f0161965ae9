// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the image-processing arithmetic the VINS-Mono front end
// delegates to OpenCV (the OpenCV sources are NOT in /root/reference; the pinned
// release is 3.3.1 via ros:kinetic-perception, docker/Dockerfile:1).  Each
// function names the reference call site it stands in for.  Parity is pinned
// against cv2 4.13.0 (the only OpenCV available offline): see
// tests/test_oracle_frontend.py and tests/golden/make_frontend_golden.py.
//
//   orc_clahe      <- cv::createCLAHE(3.0, Size(8,8))->apply   feature_tracker.cpp:87-93
//   orc_pyrdown    <- cv::pyrDown inside buildOpticalFlowPyramid feature_tracker.cpp:113
//   orc_lk         <- cv::calcOpticalFlowPyrLK(..., Size(21,21), 3) feature_tracker.cpp:113
//   orc_min_eig    <- cv::cornerMinEigenVal (inside goodFeaturesToTrack) feature_tracker.cpp:149
//   orc_gftt       <- cv::goodFeaturesToTrack(img, n, 0.01, MIN_DIST, mask) feature_tracker.cpp:149
//   orc_circle     <- cv::circle(mask, pt, MIN_DIST, 0, -1)    feature_tracker.cpp:66
//
// Build: g++ -O2 -ffp-contract=off (no FMA contraction: OpenCV's baseline is SSE3).
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>

namespace {

inline int cv_round_f(float v) { return (int)lrintf(v); }   // round-half-even (default FE mode)
inline int cv_round_d(double v) { return (int)lrint(v); }
inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
inline uint8_t sat_u8_from_float(float v) { int i = cv_round_f(v); return (uint8_t)(i < 0 ? 0 : i > 255 ? 255 : i); }

// BORDER_REFLECT_101 index (single reflection is enough for |overshoot| < len; loop for safety)
inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------
// CLAHE, clipLimit/tiles as parameters (VINS uses 3.0, 8x8).  u8 only.
// Follows OpenCV imgproc/clahe.cpp: CLAHE_CalcLut_Body + CLAHE_Interpolation_Body.
void orc_clahe(const uint8_t* src, int rows, int cols, int src_stride, double clip_limit, int tiles_x,
               int tiles_y, uint8_t* dst, int dst_stride) {
    const int hist_size = 256;
    int ext_rows = rows, ext_cols = cols;
    if (cols % tiles_x != 0 || rows % tiles_y != 0) {
        // OpenCV pads by tiles - (size % tiles) on both axes, even when one of them divides evenly.
        ext_rows = rows + (tiles_y - (rows % tiles_y));
        ext_cols = cols + (tiles_x - (cols % tiles_x));
    }
    const int tw = ext_cols / tiles_x, th = ext_rows / tiles_y;
    const int tile_area = tw * th;
    const float lut_scale = (float)(hist_size - 1) / tile_area;
    int clip = 0;
    if (clip_limit > 0.0) {
        clip = (int)(clip_limit * tile_area / hist_size);
        clip = std::max(clip, 1);
    }
    std::vector<uint8_t> lut((size_t)tiles_x * tiles_y * hist_size);
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++) {
            int hist[256];
            std::memset(hist, 0, sizeof(hist));
            for (int y = 0; y < th; y++) {
                int sy = reflect101(ty * th + y, rows);  // bottom/right padding is REFLECT_101 of src
                for (int x = 0; x < tw; x++) {
                    int sx = reflect101(tx * tw + x, cols);
                    hist[src[(size_t)sy * src_stride + sx]]++;
                }
            }
            if (clip > 0) {
                int clipped = 0;
                for (int i = 0; i < hist_size; i++)
                    if (hist[i] > clip) {
                        clipped += hist[i] - clip;
                        hist[i] = clip;
                    }
                int redist = clipped / hist_size;
                int residual = clipped - redist * hist_size;
                for (int i = 0; i < hist_size; i++) hist[i] += redist;
                if (residual != 0) {
                    int step = std::max(hist_size / residual, 1);
                    for (int i = 0; i < hist_size && residual > 0; i += step, residual--) hist[i]++;
                }
            }
            uint8_t* tl = &lut[((size_t)ty * tiles_x + tx) * hist_size];
            int sum = 0;
            for (int i = 0; i < hist_size; i++) {
                sum += hist[i];
                tl[i] = sat_u8_from_float(sum * lut_scale);
            }
        }
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    std::vector<int> ind1(cols), ind2(cols);
    std::vector<float> xa(cols), xa1(cols);
    for (int x = 0; x < cols; x++) {
        float txf = x * inv_tw - 0.5f;
        int tx1 = cv_floor_f(txf), tx2 = tx1 + 1;
        xa[x] = txf - tx1;
        xa1[x] = 1.0f - xa[x];
        tx1 = std::max(tx1, 0);
        tx2 = std::min(tx2, tiles_x - 1);
        ind1[x] = tx1 * hist_size;
        ind2[x] = tx2 * hist_size;
    }
    for (int y = 0; y < rows; y++) {
        float tyf = y * inv_th - 0.5f;
        int ty1 = cv_floor_f(tyf), ty2 = ty1 + 1;
        float ya = tyf - ty1, ya1 = 1.0f - ya;
        ty1 = std::max(ty1, 0);
        ty2 = std::min(ty2, tiles_y - 1);
        const uint8_t* p1 = &lut[(size_t)ty1 * tiles_x * hist_size];
        const uint8_t* p2 = &lut[(size_t)ty2 * tiles_x * hist_size];
        for (int x = 0; x < cols; x++) {
            int v = src[(size_t)y * src_stride + x];
            int i1 = ind1[x] + v, i2 = ind2[x] + v;
            float res = (p1[i1] * xa1[x] + p1[i2] * xa[x]) * ya1 + (p2[i1] * xa1[x] + p2[i2] * xa[x]) * ya;
            dst[(size_t)y * dst_stride + x] = sat_u8_from_float(res);
        }
    }
}

// ---------------------------------------------------------------------------
// pyrDown u8: separable [1 4 6 4 1], reflect-101, (sum+128)>>8.  OpenCV imgproc/pyramids.cpp.
void orc_pyrdown(const uint8_t* src, int rows, int cols, int src_stride, uint8_t* dst, int dst_stride) {
    const int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
    for (int dy = 0; dy < drows; dy++)
        for (int dx = 0; dx < dcols; dx++) {
            int acc = 0;
            static const int k[5] = {1, 4, 6, 4, 1};
            for (int j = -2; j <= 2; j++) {
                int sy = reflect101(2 * dy + j, rows);
                int row = 0;
                for (int i = -2; i <= 2; i++) {
                    int sx = reflect101(2 * dx + i, cols);
                    row += k[i + 2] * src[(size_t)sy * src_stride + sx];
                }
                acc += k[j + 2] * row;
            }
            dst[(size_t)dy * dst_stride + dx] = (uint8_t)((acc + 128) >> 8);
        }
}

// Number of pyramid levels calcOpticalFlowPyrLK really uses (maxLevel may be cut when a level gets
// smaller than the window; OpenCV video/lkpyramid.cpp buildOpticalFlowPyramid).
int orc_lk_num_levels(int rows, int cols, int win, int max_level) {
    int lv = 0, r = rows, c = cols;
    while (lv < max_level) {
        int nr = (r + 1) / 2, nc = (c + 1) / 2;
        if (nc <= win || nr <= win) break;
        r = nr;
        c = nc;
        lv++;
    }
    return lv;
}

}  // extern "C"

namespace {

struct Level {
    int rows, cols;
    std::vector<uint8_t> img;
    std::vector<int16_t> deriv;  // interleaved dx,dy (Scharr), only for the previous image
    inline int px(int y, int x) const {  // reflect-101 padded access (pyrBorder = BORDER_REFLECT_101)
        return img[(size_t)reflect101(y, rows) * cols + reflect101(x, cols)];
    }
    inline int dxy(int y, int x, int c) const {  // BORDER_CONSTANT 0 outside
        if (x < 0 || x >= cols || y < 0 || y >= rows) return 0;
        return deriv[((size_t)y * cols + x) * 2 + c];
    }
};

void build_pyramid(const uint8_t* img, int rows, int cols, int stride, int nlev, std::vector<Level>& pyr) {
    pyr.resize(nlev + 1);
    pyr[0].rows = rows;
    pyr[0].cols = cols;
    pyr[0].img.resize((size_t)rows * cols);
    for (int y = 0; y < rows; y++) std::memcpy(&pyr[0].img[(size_t)y * cols], img + (size_t)y * stride, cols);
    for (int l = 1; l <= nlev; l++) {
        pyr[l].rows = (pyr[l - 1].rows + 1) / 2;
        pyr[l].cols = (pyr[l - 1].cols + 1) / 2;
        pyr[l].img.resize((size_t)pyr[l].rows * pyr[l].cols);
        orc_pyrdown(pyr[l - 1].img.data(), pyr[l - 1].rows, pyr[l - 1].cols, pyr[l - 1].cols, pyr[l].img.data(),
                    pyr[l].cols);
    }
}

// calcSharrDeriv (lkpyramid.cpp): vertical [3 10 3]/[-1 0 1] then horizontal, reflect-101 on the level itself.
void scharr_deriv(Level& L) {
    const int rows = L.rows, cols = L.cols;
    L.deriv.resize((size_t)rows * cols * 2);
    std::vector<int> t0(cols + 2), t1(cols + 2);
    for (int y = 0; y < rows; y++) {
        const uint8_t* r0 = &L.img[(size_t)(y > 0 ? y - 1 : rows > 1 ? 1 : 0) * cols];
        const uint8_t* r1 = &L.img[(size_t)y * cols];
        const uint8_t* r2 = &L.img[(size_t)(y < rows - 1 ? y + 1 : rows > 1 ? rows - 2 : 0) * cols];
        for (int x = 0; x < cols; x++) {
            t0[x + 1] = (int16_t)((r0[x] + r2[x]) * 3 + r1[x] * 10);
            t1[x + 1] = (int16_t)(r2[x] - r0[x]);
        }
        int x0 = cols > 1 ? 1 : 0, x1 = cols > 1 ? cols - 2 : 0;
        t0[0] = t0[x0 + 1];
        t0[cols + 1] = t0[x1 + 1];
        t1[0] = t1[x0 + 1];
        t1[cols + 1] = t1[x1 + 1];
        for (int x = 0; x < cols; x++) {
            L.deriv[((size_t)y * cols + x) * 2 + 0] = (int16_t)(t0[x + 2] - t0[x]);
            L.deriv[((size_t)y * cols + x) * 2 + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
}

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------
// Pyramidal LK exactly as OpenCV's LKTrackerInvoker (video/lkpyramid.cpp), flags = 0,
// criteria = COUNT+EPS (30, 0.01), minEigThreshold 1e-4.  The window sums A11/A12/A22/b1/b2 are sums
// of integer products accumulated in float exactly the way OpenCV's x86 (CV_SIMD128) build does it: 16 pixels
// of every window row through four float lanes, the remaining 5 through a scalar float, lanes combined as
// (l0 + l2) + (l1 + l3) at the end.  Against cv2 4.13 this restatement is bit-exact (coordinates and status;
// tests/test_oracle_frontend.py).  orc_lk_set_simd_sums(0) switches to exact int64 sums (the accumulator type of
// OpenCV's NEON build), which differ from the x86 results by <= ~1e-4 px.
// 1 (default): window sums accumulated like OpenCV's x86 SIMD path (bit-exact against cv2 4.13, verified in
// tests/test_oracle_frontend.py); 0: exact int64 sums (OpenCV's NEON accumulator type), kept for comparison.
static int simd_sums = 1;
void orc_lk_set_simd_sums(int on) { simd_sums = on != 0; }

void orc_lk(const uint8_t* prev, const uint8_t* next, int rows, int cols, int stride, const float* prev_pts, int n,
            int win, int max_level, int max_iter, double eps, double min_eig_thr, float* next_pts,
            uint8_t* status) {
    const int nlev = orc_lk_num_levels(rows, cols, win, max_level);
    std::vector<Level> P, N;
    build_pyramid(prev, rows, cols, stride, nlev, P);
    build_pyramid(next, rows, cols, stride, nlev, N);
    for (auto& L : P) scharr_deriv(L);
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const double epsilon = eps * eps;  // OpenCV squares criteria.epsilon (stays double)
    std::vector<int16_t> Iw((size_t)win * win), dIw((size_t)win * win * 2);
    for (int i = 0; i < n; i++) status[i] = 1;
    for (int level = nlev; level >= 0; level--) {
        const Level& I = P[level];
        const Level& J = N[level];
        for (int p = 0; p < n; p++) {
            float ppx = prev_pts[2 * p] * (float)(1. / (1 << level));
            float ppy = prev_pts[2 * p + 1] * (float)(1. / (1 << level));
            float nx, ny;
            if (level == nlev) {
                nx = ppx;
                ny = ppy;
            } else {
                nx = next_pts[2 * p] * 2.f;
                ny = next_pts[2 * p + 1] * 2.f;
            }
            next_pts[2 * p] = nx;
            next_pts[2 * p + 1] = ny;
            ppx -= half;
            ppy -= half;
            int ipx = cv_floor_f(ppx), ipy = cv_floor_f(ppy);
            if (ipx < -win || ipx >= I.cols || ipy < -win || ipy >= I.rows) {
                if (level == 0) status[p] = 0;
                continue;
            }
            float a = ppx - ipx, b = ppy - ipy;
            const int W_BITS = 14;
            int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
            int iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
            int iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
            int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            // Window sums in OpenCV's own accumulation structure (video/lkpyramid.cpp, CV_SIMD128 path of the x86 builds):
            // per row the first 16 pixels go through four float lanes (lane j takes pixels j, 4+j, 8+j, 12+j), the last
            // five through a scalar float accumulator; at the end total = scalar + ((l0 + l2) + (l1 + l3)).
            float qA11[4] = {0, 0, 0, 0}, qA12[4] = {0, 0, 0, 0}, qA22[4] = {0, 0, 0, 0}, sA11 = 0, sA12 = 0, sA22 = 0;
            const int nsimd = simd_sums ? (win / 8) * 8 : 0;
            int64_t iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    int yy = y + ipy, xx = x + ipx;
                    int ival = descale(I.px(yy, xx) * iw00 + I.px(yy, xx + 1) * iw01 + I.px(yy + 1, xx) * iw10 +
                                           I.px(yy + 1, xx + 1) * iw11,
                                       W_BITS - 5);
                    int ixval = descale(I.dxy(yy, xx, 0) * iw00 + I.dxy(yy, xx + 1, 0) * iw01 +
                                            I.dxy(yy + 1, xx, 0) * iw10 + I.dxy(yy + 1, xx + 1, 0) * iw11,
                                        W_BITS);
                    int iyval = descale(I.dxy(yy, xx, 1) * iw00 + I.dxy(yy, xx + 1, 1) * iw01 +
                                            I.dxy(yy + 1, xx, 1) * iw10 + I.dxy(yy + 1, xx + 1, 1) * iw11,
                                        W_BITS);
                    Iw[y * win + x] = (int16_t)ival;
                    dIw[(y * win + x) * 2] = (int16_t)ixval;
                    dIw[(y * win + x) * 2 + 1] = (int16_t)iyval;
                    iA11 += (int64_t)ixval * ixval;
                    iA12 += (int64_t)ixval * iyval;
                    iA22 += (int64_t)iyval * iyval;
                    if (x < nsimd) {
                        const float fx = (float)ixval, fy = (float)iyval;
                        qA22[x & 3] += fy * fy;
                        qA12[x & 3] += fx * fy;
                        qA11[x & 3] += fx * fx;
                    } else {
                        sA11 += (float)(ixval * ixval);
                        sA12 += (float)(ixval * iyval);
                        sA22 += (float)(iyval * iyval);
                    }
                }
            float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
            if (simd_sums) {
                A11 = (sA11 + ((qA11[0] + qA11[2]) + (qA11[1] + qA11[3]))) * FLT_SCALE;
                A12 = (sA12 + ((qA12[0] + qA12[2]) + (qA12[1] + qA12[3]))) * FLT_SCALE;
                A22 = (sA22 + ((qA22[0] + qA22[2]) + (qA22[1] + qA22[3]))) * FLT_SCALE;
            }
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < (float)min_eig_thr || D < FLT_EPSILON) {
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= half;
            ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_iter; j++) {
                int inx = cv_floor_f(nx), iny = cv_floor_f(ny);
                if (inx < -win || inx >= J.cols || iny < -win || iny >= J.rows) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                a = nx - inx;
                b = ny - iny;
                iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
                iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                int64_t ib1 = 0, ib2 = 0;
                // mismatch vector: per row and block of 8 pixels, lane pairs (k, k+4) are summed as integers
                // (v_dotprod), converted to float and added to the float lanes qb0 (k = 0, 1) / qb1 (k = 2, 3)
                float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0}, sb1 = 0, sb2 = 0;
                int dbuf[8], xbuf[8], ybuf[8];
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        int yy = y + iny, xx = x + inx;
                        int diff = descale(J.px(yy, xx) * iw00 + J.px(yy, xx + 1) * iw01 + J.px(yy + 1, xx) * iw10 +
                                               J.px(yy + 1, xx + 1) * iw11,
                                           W_BITS - 5) -
                                   Iw[y * win + x];
                        const int ixv = dIw[(y * win + x) * 2], iyv = dIw[(y * win + x) * 2 + 1];
                        ib1 += (int64_t)diff * ixv;
                        ib2 += (int64_t)diff * iyv;
                        if (x < nsimd) {
                            dbuf[x & 7] = diff;
                            xbuf[x & 7] = ixv;
                            ybuf[x & 7] = iyv;
                            if ((x & 7) == 7) {
                                qb0[0] += (float)(dbuf[0] * xbuf[0] + dbuf[4] * xbuf[4]);
                                qb0[1] += (float)(dbuf[0] * ybuf[0] + dbuf[4] * ybuf[4]);
                                qb0[2] += (float)(dbuf[1] * xbuf[1] + dbuf[5] * xbuf[5]);
                                qb0[3] += (float)(dbuf[1] * ybuf[1] + dbuf[5] * ybuf[5]);
                                qb1[0] += (float)(dbuf[2] * xbuf[2] + dbuf[6] * xbuf[6]);
                                qb1[1] += (float)(dbuf[2] * ybuf[2] + dbuf[6] * ybuf[6]);
                                qb1[2] += (float)(dbuf[3] * xbuf[3] + dbuf[7] * xbuf[7]);
                                qb1[3] += (float)(dbuf[3] * ybuf[3] + dbuf[7] * ybuf[7]);
                            }
                        } else {
                            sb1 += (float)(diff * ixv);
                            sb2 += (float)(diff * iyv);
                        }
                    }
                float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
                if (simd_sums) {
                    const float q0 = qb0[0] + qb1[0], q1 = qb0[1] + qb1[1], q2 = qb0[2] + qb1[2], q3 = qb0[3] + qb1[3];
                    b1 = (sb1 + (q0 + q2)) * FLT_SCALE;
                    b2 = (sb2 + (q1 + q3)) * FLT_SCALE;
                }
                float dx = (float)((A12 * b2 - A22 * b1) * D);
                float dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx;
                ny += dy;
                next_pts[2 * p] = nx + half;
                next_pts[2 * p + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= epsilon) break;  // Point2f::ddot -> double
                if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) {
                    next_pts[2 * p] -= dx * 0.5f;
                    next_pts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx;
                pdy = dy;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// cornerMinEigenVal(blockSize 3, ksize 3) on u8: Sobel (scale 1/3060 folded into the smoothing
// kernel) -> products -> 3x3 un-normalised box sum with double accumulators (sumType CV_64F for float
// input, imgproc/box_filter.dispatch.cpp createBoxFilter) -> (a+c) - sqrt((a-c)^2 + b^2).
void orc_min_eig(const uint8_t* img, int rows, int cols, int stride, float* eig) {
    const double scale_d = 1.0 / ((double)(1 << 2) * 3 * 255.0);
    const float k1 = (float)scale_d;         // smoothing kernel [1 2 1]*scale, stored as CV_32F
    const float k2 = (float)(2.0 * scale_d);
    std::vector<float> dx((size_t)rows * cols), dy((size_t)rows * cols);
    auto P = [&](int y, int x) -> float { return (float)img[(size_t)reflect101(y, rows) * stride + reflect101(x, cols)]; };
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            // Sobel dx: row filter [-1 0 1] (exact), column filter symmetric [k1 k2 k1]:
            //   SymmColumnFilter: s = k2*S0 + k1*(S-1 + S+1)
            float rm = P(y - 1, x + 1) - P(y - 1, x - 1);
            float r0 = P(y, x + 1) - P(y, x - 1);
            float rp = P(y + 1, x + 1) - P(y + 1, x - 1);
            dx[(size_t)y * cols + x] = k2 * r0 + k1 * (rm + rp);
            // Sobel dy: generic RowFilter<uchar,float> runs the taps in order: (k1*S-1 + k2*S0) + k1*S+1;
            // column filter [-1 0 1] is an exact difference.
            float sm = (k1 * P(y - 1, x - 1) + k2 * P(y - 1, x)) + k1 * P(y - 1, x + 1);
            float sp = (k1 * P(y + 1, x - 1) + k2 * P(y + 1, x)) + k1 * P(y + 1, x + 1);
            dy[(size_t)y * cols + x] = sp - sm;
        }
    auto DX = [&](int y, int x) -> float { return dx[(size_t)reflect101(y, rows) * cols + reflect101(x, cols)]; };
    auto DY = [&](int y, int x) -> float { return dy[(size_t)reflect101(y, rows) * cols + reflect101(x, cols)]; };
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            double sa = 0, sb = 0, sc = 0;
            for (int j = -1; j <= 1; j++) {
                double ra = 0, rb = 0, rc = 0;
                for (int i = -1; i <= 1; i++) {
                    float gx = DX(y + j, x + i), gy = DY(y + j, x + i);
                    ra += (double)(gx * gx);
                    rb += (double)(gx * gy);
                    rc += (double)(gy * gy);
                }
                sa += ra;
                sb += rb;
                sc += rc;
            }
            float a = (float)sa * 0.5f, b = (float)sb, c = (float)sc * 0.5f;
            eig[(size_t)y * cols + x] = (float)((a + c) - std::sqrt((a - c) * (a - c) + b * b));
        }
}

// goodFeaturesToTrack(image, corners, maxCorners, qualityLevel, minDistance, mask) with blockSize 3,
// useHarris false (imgproc/featureselect.cpp).  mask may be NULL.  Returns the number of corners;
// corners are integer pixel positions as float (x,y).
int orc_gftt(const uint8_t* img, int rows, int cols, int stride, const uint8_t* mask, int mask_stride,
             int max_corners, double quality, double min_distance, float* corners, int* n_candidates) {
    std::vector<float> eig((size_t)rows * cols);
    orc_min_eig(img, rows, cols, stride, eig.data());
    float max_val = 0;  // minMaxLoc(eig, 0, &maxVal, 0, 0, mask)
    bool any = false;
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++)
            if (!mask || mask[(size_t)y * mask_stride + x]) {
                float v = eig[(size_t)y * cols + x];
                if (!any || v > max_val) max_val = v, any = true;
            }
    if (!any) max_val = 0;
    const float thr = (float)((double)max_val * quality);  // threshold(eig, eig, maxVal*q, 0, THRESH_TOZERO)
    for (auto& v : eig)
        if (!(v > thr)) v = 0;
    std::vector<int> cand;  // linear offsets
    for (int y = 1; y < rows - 1; y++)
        for (int x = 1; x < cols - 1; x++) {
            float v = eig[(size_t)y * cols + x];
            if (v == 0) continue;
            if (mask && !mask[(size_t)y * mask_stride + x]) continue;
            float m = v;  // dilate 3x3 (all neighbours are inside the image for interior pixels)
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) m = std::max(m, eig[(size_t)(y + j) * cols + (x + i)]);
            if (v == m) cand.push_back(y * cols + x);
        }
    if (n_candidates) *n_candidates = (int)cand.size();
    std::sort(cand.begin(), cand.end(), [&](int a, int b) {
        float va = eig[a], vb = eig[b];
        return va > vb ? true : va < vb ? false : a > b;  // greaterThanPtr: ties by address, descending
    });
    int ncorners = 0;
    if (min_distance >= 1) {
        const int cell = cv_round_d(min_distance);
        const int gw = (cols + cell - 1) / cell, gh = (rows + cell - 1) / cell;
        std::vector<std::vector<float>> grid((size_t)gw * gh);
        const float md2 = (float)(min_distance * min_distance);
        for (size_t i = 0; i < cand.size(); i++) {
            int y = cand[i] / cols, x = cand[i] % cols;
            int xc = x / cell, yc = y / cell;
            int x1 = std::max(0, xc - 1), y1 = std::max(0, yc - 1);
            int x2 = std::min(gw - 1, xc + 1), y2 = std::min(gh - 1, yc + 1);
            bool good = true;
            for (int yy = y1; yy <= y2 && good; yy++)
                for (int xx = x1; xx <= x2 && good; xx++) {
                    const std::vector<float>& m = grid[(size_t)yy * gw + xx];
                    for (size_t j = 0; j < m.size(); j += 2) {
                        float ddx = x - m[j], ddy = y - m[j + 1];
                        if (ddx * ddx + ddy * ddy < md2) {
                            good = false;
                            break;
                        }
                    }
                }
            if (good) {
                grid[(size_t)yc * gw + xc].push_back((float)x);
                grid[(size_t)yc * gw + xc].push_back((float)y);
                corners[2 * ncorners] = (float)x;
                corners[2 * ncorners + 1] = (float)y;
                ++ncorners;
                if (max_corners > 0 && ncorners == max_corners) break;
            }
        }
    } else {
        for (size_t i = 0; i < cand.size(); i++) {
            corners[2 * ncorners] = (float)(cand[i] % cols);
            corners[2 * ncorners + 1] = (float)(cand[i] / cols);
            ++ncorners;
            if (max_corners > 0 && ncorners == max_corners) break;
        }
    }
    return ncorners;
}

// ---------------------------------------------------------------------------
// cv::circle(img, Point(cx,cy), radius, color, FILLED) on a u8 image (imgproc/drawing.cpp Circle(),
// fill branch), clipped to the image.
void orc_circle(uint8_t* img, int rows, int cols, int stride, int cx, int cy, int radius, uint8_t color) {
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    auto hline = [&](int y, int xa, int xb) {
        if (y < 0 || y >= rows) return;
        xa = std::max(xa, 0);
        xb = std::min(xb, cols - 1);
        for (int x = xa; x <= xb; x++) img[(size_t)y * stride + x] = color;
    };
    while (dx >= dy) {
        int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        hline(y11, x11, x12);
        hline(y12, x11, x12);
        hline(y21, x21, x22);
        hline(y22, x21, x22);
        dy++;
        err += plus;
        plus += 2;
        int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

}  // extern "C"
