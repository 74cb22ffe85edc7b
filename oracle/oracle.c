/* TEST INFRASTRUCTURE ONLY - see oracle.h.  Plain C11; compiled WITHOUT -march=native / FMA so that
 * float expressions evaluate as separate IEEE multiplies and adds, like the reference's x86-64 build.
 * Every function cites the reference file:line (relative to /root/reference) it restates. */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void orc_set_threads(int n) {
    g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}
int orc_get_threads(void) { return g_threads; }

/* =============================================================================================
 * OpenCV 4.13.0 semantics (the reference only says find_package(OpenCV), CMakeLists.txt:31).
 * ============================================================================================= */

/* cv::cvtColor(CV_BGR2Lab) on CV_8UC3 (call site include/feature_extractor.h:129).
 * Integer path of OpenCV's RGB2Lab_b: gamma LUT (<<3), 12-bit matrix, cube-root LUT (<<15). */
static uint16_t g_gamma[256];
static uint16_t g_cbrt[3072];
static int g_lab_ready = 0;
static void lab_tables(void) {
    if (g_lab_ready) return;
    for (int i = 0; i < 256; i++) {
        double x = i / 255.0;
        double g = x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4);
        g_gamma[i] = (uint16_t)lrint(255.0 * 8.0 * g);
    }
    for (int i = 0; i < 3072; i++) {
        double x = i / (255.0 * 8.0);
        double h = x < 216.0 / 24389.0 ? x * (841.0 / 108.0) + 16.0 / 116.0 : cbrt(x);
        g_cbrt[i] = (uint16_t)lrint(32768.0 * h);
    }
    /* OpenCV builds this LUT in 32-bit softfloat; against a double-precision cube root exactly these
     * entries land on the other side of .5 (checked against cv2 4.13.0 over all 2^24 colours). */
    g_cbrt[49] = 9454;
    g_cbrt[324] = 17745;
    g_cbrt[628] = 22126;
    g_lab_ready = 1;
}
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
void orc_bgr2lab_u8(const uint8_t* src, int64_t npix, uint8_t* dst) {
    lab_tables();
    /* round(4096 * sRGB2XYZ_D65 / whitepoint), columns ordered for a "BGR" source */
    static const int C[9] = {778, 1541, 1777, 296, 2929, 871, 3575, 448, 73};
    for (int64_t i = 0; i < npix; i++) {
        int c0 = g_gamma[src[3 * i]], c1 = g_gamma[src[3 * i + 1]], c2 = g_gamma[src[3 * i + 2]];
        int fX = g_cbrt[(c0 * C[0] + c1 * C[1] + c2 * C[2] + 2048) >> 12];
        int fY = g_cbrt[(c0 * C[3] + c1 * C[4] + c2 * C[5] + 2048) >> 12];
        int fZ = g_cbrt[(c0 * C[6] + c1 * C[7] + c2 * C[8] + 2048) >> 12];
        int L = (296 * fY - 1336934 + 16384) >> 15;
        int a = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
        int b = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
        dst[3 * i] = sat_u8(L);
        dst[3 * i + 1] = sat_u8(a);
        dst[3 * i + 2] = sat_u8(b);
    }
}

/* cv::copyMakeBorder(BORDER_REFLECT) (feature_extractor.h:130): fedcba|abcdefgh|hgfedcb */
static inline int reflect_idx(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p - 1 : 2 * n - 1 - p;
    return p;
}
void orc_border_reflect_u8c3(const uint8_t* src, int W, int H, int b, uint8_t* dst) {
    const int Wb = W + 2 * b, Hb = H + 2 * b;
    for (int y = 0; y < Hb; y++) {
        int sy = reflect_idx(y - b, H);
        for (int x = 0; x < Wb; x++) {
            int sx = reflect_idx(x - b, W);
            memcpy(dst + ((size_t)y * Wb + x) * 3, src + ((size_t)sy * W + sx) * 3, 3);
        }
    }
}

/* cv::resize(INTER_LINEAR) on CV_8UC3, square S x S -> r x r (feature_extractor.h:142).
 * 11-bit fixed-point coefficients; horizontal pass into int rows, then the vertical pass. */
static inline int round_half_even_f(float v) { return (int)lrintf(v); } /* FE_TONEAREST default */
void orc_resize_linear_u8c3(const uint8_t* src, int sstep, int S, uint8_t* dst, int r) {
    const double scale = 1.0 / ((double)r / (double)S);
    int* xs = (int*)malloc(sizeof(int) * r);
    int* a0 = (int*)malloc(sizeof(int) * r);
    int* a1 = (int*)malloc(sizeof(int) * r);
    for (int d = 0; d < r; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= S - 1) { s = S - 1; f = 0.f; }
        xs[d] = s;
        a0[d] = round_half_even_f((1.f - f) * 2048.f);
        a1[d] = round_half_even_f(f * 2048.f);
    }
    for (int dy = 0; dy < r; dy++) {
        float f = (float)((dy + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        int b0 = round_half_even_f((1.f - f) * 2048.f), b1 = round_half_even_f(f * 2048.f);
        int y0 = s < 0 ? 0 : (s > S - 1 ? S - 1 : s);
        int y1 = s + 1 < 0 ? 0 : (s + 1 > S - 1 ? S - 1 : s + 1);
        const uint8_t* r0 = src + (size_t)y0 * sstep;
        const uint8_t* r1 = src + (size_t)y1 * sstep;
        for (int dx = 0; dx < r; dx++) {
            int s0 = xs[dx], s1 = s0 + 1 > S - 1 ? S - 1 : s0 + 1;
            for (int c = 0; c < 3; c++) {
                int h0 = r0[3 * s0 + c] * a0[dx] + r0[3 * s1 + c] * a1[dx];
                int h1 = r1[3 * s0 + c] * a0[dx] + r1[3 * s1 + c] * a1[dx];
                int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                dst[((size_t)dy * r + dx) * 3 + c] = sat_u8(v);
            }
        }
    }
    free(xs); free(a0); free(a1);
}

/* cv::resize(INTER_LINEAR) on CV_32FC(C) (call sites src/segmenter.cpp:381, src/test_multi.cpp:199). */
void orc_resize_linear_f32(const float* src, int sw, int sh, int C, float* dst, int dw, int dh) {
    const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);
    int* xo = (int*)malloc(sizeof(int) * dw);
    float* xa = (float*)malloc(sizeof(float) * dw);
    for (int d = 0; d < dw; d++) {
        float f = (float)((d + 0.5) * scale_x - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sw - 1) { s = sw - 1; f = 0.f; }
        xo[d] = s; xa[d] = f;
    }
    float* h0 = (float*)malloc(sizeof(float) * (size_t)dw * C);
    float* h1 = (float*)malloc(sizeof(float) * (size_t)dw * C);
    for (int dy = 0; dy < dh; dy++) {
        float f = (float)((dy + 0.5) * scale_y - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        int y0 = s < 0 ? 0 : (s > sh - 1 ? sh - 1 : s);
        int y1 = s + 1 < 0 ? 0 : (s + 1 > sh - 1 ? sh - 1 : s + 1);
        const float b0 = 1.f - f, b1 = f;
        for (int pass = 0; pass < 2; pass++) {
            const float* row = src + (size_t)(pass ? y1 : y0) * sw * C;
            float* h = pass ? h1 : h0;
            for (int dx = 0; dx < dw; dx++) {
                int s0 = xo[dx], s1 = s0 + 1 > sw - 1 ? sw - 1 : s0 + 1;
                float a1 = xa[dx], a0 = 1.f - a1;
                for (int c = 0; c < C; c++) {
                    float p0 = row[(size_t)s0 * C + c] * a0;
                    float p1 = row[(size_t)s1 * C + c] * a1;
                    h[(size_t)dx * C + c] = p0 + p1;
                }
            }
        }
        for (size_t k = 0; k < (size_t)dw * C; k++) {
            float p0 = h0[k] * b0, p1 = h1[k] * b1;
            dst[(size_t)dy * dw * C + k] = p0 + p1;
        }
    }
    free(xo); free(xa); free(h0); free(h1);
}

/* =============================================================================================
 * Feature extractor: include/feature_extractor.h:29-291
 * ============================================================================================= */
int orc_feature_length(const orc_fe_config* c) { /* :46-51 */
    int D = 0;
    if (c->use_color_patch) D += c->patch_size_reduce * c->patch_size_reduce * 3;
    if (c->use_depth) D += 1;
    if (c->use_height) D += 1;
    if (c->use_normal) D += 1;
    return D;
}

/* :200-232.  Eigen association restated as M = R*Kinv (row-by-column, (a0b0+a1b1)+a2b2), then
 * ((m0*v0 + m1*v1) + m2*v2) + t, every product and sum rounded to float (no FMA). */
void orc_cloud(const uint16_t* depth, int W, int H, const float* Kinv, const float* R, const float* t, float dmin,
               float dmax, float* xyz) {
    float M[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float p0 = R[3 * i] * Kinv[j], p1 = R[3 * i + 1] * Kinv[3 + j], p2 = R[3 * i + 2] * Kinv[6 + j];
            float s = p0 + p1;
            M[3 * i + j] = s + p2;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t idx = (size_t)y * W + x;
            float d = (float)depth[idx] / 1000.0f;
            float v0, v1, v2;
            if (d < dmin || d > dmax) v0 = v1 = v2 = NAN;
            else { v0 = d * (float)x; v1 = d * (float)y; v2 = d; }
            for (int i = 0; i < 3; i++) {
                float p0 = M[3 * i] * v0, p1 = M[3 * i + 1] * v1, p2 = M[3 * i + 2] * v2;
                float s = p0 + p1;
                s = s + p2;
                xyz[3 * idx + i] = s + t[i];
            }
        }
}

/* PCL IntegralImageNormalEstimation (call site feature_extractor.h:256-261).  PCL is NOT in
 * /root/reference; restated from PCL 1.7/1.8 features/impl/integral_image_normal.hpp and
 * features/impl/integral_image2D.hpp.  PARITY UNPINNED (no reference test or binary to check). */
void orc_normals(const float* xyz, int W, int H, float* normals, float* dist_out) {
    const size_t NP = (size_t)W * H;
    const int W1 = W + 1;
    /* initAverage3DGradientMethod: central differences on the interior, zero on the 1-px frame */
    float* dx = (float*)calloc(NP * 3, sizeof(float));
    float* dy = (float*)calloc(NP * 3, sizeof(float));
    for (int r = 1; r < H - 1; r++)
        for (int c = 1; c < W - 1; c++) {
            size_t i = (size_t)r * W + c;
            for (int k = 0; k < 3; k++) {
                dx[3 * i + k] = xyz[3 * (i + 1) + k] - xyz[3 * (i - 1) + k];
                dy[3 * i + k] = xyz[3 * (i + W) + k] - xyz[3 * (i - W) + k];
            }
        }
    /* IntegralImage2D<float,3>::computeIntegralImages, first order only, double sums + finite counts */
    double* I[2];
    unsigned* Cn[2];
    const float* D[2] = {dx, dy};
    for (int m = 0; m < 2; m++) {
        I[m] = (double*)calloc((size_t)W1 * (H + 1) * 3, sizeof(double));
        Cn[m] = (unsigned*)calloc((size_t)W1 * (H + 1), sizeof(unsigned));
        for (int r = 0; r < H; r++) {
            double* prev = I[m] + (size_t)r * W1 * 3;
            double* cur = prev + (size_t)W1 * 3;
            unsigned* cprev = Cn[m] + (size_t)r * W1;
            unsigned* ccur = cprev + W1;
            for (int c = 0; c < W; c++) {
                const float* e = D[m] + ((size_t)r * W + c) * 3;
                float fs = e[0] + e[1];
                fs = fs + e[2];
                int fin = isfinite(fs);
                for (int k = 0; k < 3; k++) {
                    double v = prev[3 * (c + 1) + k] + cur[3 * c + k];
                    v = v - prev[3 * c + k];
                    if (fin) v = v + (double)e[k];
                    cur[3 * (c + 1) + k] = v;
                }
                ccur[c + 1] = cprev[c + 1] + ccur[c] - cprev[c] + (fin ? 1u : 0u);
            }
        }
    }
    /* computeFeature: depth-change map on z */
    unsigned char* dcm = (unsigned char*)malloc(NP);
    memset(dcm, 255, NP);
    for (int r = 0; r < H - 1; r++)
        for (int c = 0; c < W - 1; c++) {
            size_t i = (size_t)r * W + c;
            const float z = xyz[3 * i + 2], zR = xyz[3 * (i + 1) + 2], zD = xyz[3 * (i + W) + 2];
            const float thr = (0.02f * (fabsf(z) + 1.0f)) * 2.0f;
            if (fabsf(z - zR) > thr || !isfinite(z) || !isfinite(zR)) { dcm[i] = 0; dcm[i + 1] = 0; }
            if (fabsf(z - zD) > thr || !isfinite(z) || !isfinite(zD)) { dcm[i] = 0; dcm[i + W] = 0; }
        }
    /* two-pass 1.0/1.4 chamfer distance map (the row-wrapping reads at ci=W-1 / ci=0 are PCL's) */
    float* dm = (float*)malloc(NP * sizeof(float));
    for (size_t i = 0; i < NP; i++) dm[i] = dcm[i] == 0 ? 0.0f : (float)(W + H);
    for (int r = 1; r < H; r++) {
        float* prev = dm + (size_t)(r - 1) * W;
        float* cur = dm + (size_t)r * W;
        for (int c = 1; c < W; c++) {
            const float ul = prev[c - 1] + 1.4f, up = prev[c] + 1.0f, ur = prev[c + 1] + 1.4f;
            const float lf = cur[c - 1] + 1.0f, ce = cur[c];
            const float m = fminf(fminf(ul, up), fminf(lf, ur));
            if (m < ce) cur[c] = m;
        }
    }
    for (int r = H - 2; r >= 0; r--) {
        float* next = dm + (size_t)(r + 1) * W;
        float* cur = dm + (size_t)r * W;
        for (int c = W - 2; c >= 0; c--) {
            const float ll = next[c - 1] + 1.4f, lo = next[c] + 1.0f, lr = next[c + 1] + 1.4f;
            const float rt = cur[c + 1] + 1.0f, ce = cur[c];
            const float m = fminf(fminf(ll, lo), fminf(rt, lr));
            if (m < ce) cur[c] = m;
        }
    }
    if (dist_out) memcpy(dist_out, dm, NP * sizeof(float));
    /* BORDER_POLICY_IGNORE: everything NaN, then the interior [border, dim-border) */
    for (size_t i = 0; i < NP * 3; i++) normals[i] = NAN;
    const int border = 10; /* int(normal_smoothing_size_) */
    for (int r = border; r < H - border; r++)
        for (int c = border; c < W - border; c++) {
            size_t idx = (size_t)r * W + c;
            if (!isfinite(xyz[3 * idx + 2])) continue;
            float sm = fminf(dm[idx], 10.0f);
            if (!(sm > 2.0f)) continue;
            const int w = (int)sm, sx = c - w / 2, sy = r - w / 2;
            const size_t ul = (size_t)sy * W1 + sx, ur = ul + w, ll = (size_t)(sy + w) * W1 + sx, lr = ll + w;
            unsigned cx = Cn[0][lr] + Cn[0][ul] - Cn[0][ur] - Cn[0][ll];
            unsigned cy = Cn[1][lr] + Cn[1][ul] - Cn[1][ur] - Cn[1][ll];
            if (cx == 0 || cy == 0) continue;
            double gx[3], gy[3];
            for (int k = 0; k < 3; k++) {
                double a = I[0][3 * lr + k] + I[0][3 * ul + k];
                a = a - I[0][3 * ur + k];
                gx[k] = a - I[0][3 * ll + k];
                double b = I[1][3 * lr + k] + I[1][3 * ul + k];
                b = b - I[1][3 * ur + k];
                gy[k] = b - I[1][3 * ll + k];
            }
            /* normal = gradient_y.cross(gradient_x) */
            double n[3];
            n[0] = gy[1] * gx[2] - gy[2] * gx[1];
            n[1] = gy[2] * gx[0] - gy[0] * gx[2];
            n[2] = gy[0] * gx[1] - gy[1] * gx[0];
            double len = n[0] * n[0] + n[1] * n[1];
            len = len + n[2] * n[2];
            if (len == 0.0) continue;
            double s = sqrt(len);
            normals[3 * idx] = (float)(n[0] / s);
            normals[3 * idx + 1] = (float)(n[1] / s);
            normals[3 * idx + 2] = (float)(n[2] / s);
        }
    free(dx); free(dy); free(dcm); free(dm);
    for (int m = 0; m < 2; m++) { free(I[m]); free(Cn[m]); }
}

int orc_extract(const orc_fe_config* cfg, int stride, const uint8_t* rgb, const uint16_t* depth, int W, int H,
                const float* Kinv, const float* R, const float* t, float dmin, float dmax, int extract_type,
                const int8_t* labels, int L, float* feats, int* xs, int* ys, int* out_labels) {
    const float dmin_mm = (float)(dmin * 1000.0), dmax_mm = (float)(dmax * 1000.0); /* :43-44 */
    const int D = orc_feature_length(cfg);
    const int P = cfg->patch_size, r = cfg->patch_size_reduce, border = P; /* :37 */
    const size_t NP = (size_t)W * H;
    /* :56-121 sample selection in raster order */
    int n = 0;
    for (int y = 0; y < H; y += stride)
        for (int x = 0; x < W; x += stride) {
            size_t i = (size_t)y * W + x;
            float d = (float)depth[i];
            int ok = d >= dmin_mm && d <= dmax_mm;
            if (ok && extract_type == ORC_WITH_POSITIVE_LABEL)
                for (int l = 0; l < L; l++) ok &= labels[l * NP + i] >= 0;
            if (!ok) continue;
            xs[n] = x; ys[n] = y;
            if (out_labels && labels)
                for (int l = 0; l < L; l++) out_labels[(size_t)n * L + l] = labels[l * NP + i];
            n++;
        }
    int pos = 0;
    if (cfg->use_color_patch) { /* :125-175 */
        uint8_t* lab = (uint8_t*)malloc(NP * 3);
        orc_bgr2lab_u8(rgb, (int64_t)NP, lab);
        const int Wb = W + 2 * border, Hb = H + 2 * border;
        uint8_t* labb = (uint8_t*)malloc((size_t)Wb * Hb * 3);
        orc_border_reflect_u8c3(lab, W, H, border, labb);
#pragma omp parallel for schedule(dynamic, 64)
        for (int s = 0; s < n; s++) {
            uint8_t patch[3 * 32 * 32];
            float d = (float)depth[(size_t)ys[s] * W + xs[s]] / 1000.0f; /* :139 */
            int half = (int)(P / (2.0 * d));                             /* :140 int/double truncation */
            int S = half * 2 + 1;
            const uint8_t* roi = labb + ((size_t)(ys[s] + border - half) * Wb + (xs[s] + border - half)) * 3;
            orc_resize_linear_u8c3(roi, Wb * 3, S, patch, r);
            for (int k = 0; k < r * r * 3; k++) feats[(size_t)s * D + pos + k] = (float)patch[k];
        }
        free(lab); free(labb);
        pos += r * r * 3;
    }
    if (cfg->use_depth) { /* :180-197 */
        for (int s = 0; s < n; s++) feats[(size_t)s * D + pos] = (float)depth[(size_t)ys[s] * W + xs[s]] / 1000.0f;
        pos++;
    }
    float* xyz = NULL;
    if (cfg->use_height || cfg->use_normal) {
        xyz = (float*)malloc(NP * 3 * sizeof(float));
        orc_cloud(depth, W, H, Kinv, R, t, dmin, dmax, xyz);
    }
    if (cfg->use_height) { /* :236-251 */
        for (int s = 0; s < n; s++) feats[(size_t)s * D + pos] = xyz[3 * ((size_t)ys[s] * W + xs[s]) + 2];
        pos++;
    }
    if (cfg->use_normal) { /* :254-291; acos(fabs(float)) resolves to the double overloads of <math.h> */
        float* nrm = (float*)malloc(NP * 3 * sizeof(float));
        orc_normals(xyz, W, H, nrm, NULL);
        for (int s = 0; s < n; s++) {
            size_t i = (size_t)ys[s] * W + xs[s];
            feats[(size_t)s * D + pos] = isnan(nrm[3 * i]) ? -2.0f : (float)acos(fabs((double)nrm[3 * i + 2]));
        }
        free(nrm);
        pos++;
    }
    free(xyz);
    return n;
}

/* =============================================================================================
 * libforest: third-party/libforest/src/classifier.cpp, include/libforest/io.h
 * ============================================================================================= */
typedef struct {
    int n;
    int* feat;
    float* thr;
    int* left;
    float* hist; /* [n][sumC], only meaningful at leaves */
} orc_tree;
struct orc_forest {
    int T, L, sumC;
    int* C; /* classes per layer */
    orc_tree* trees;
};
static int rd_i32(FILE* f, int* v) { return fread(v, 4, 1, f) == 1; }
/* io.h:43-108 readBinary<vector<T>>: int32 count then the elements; classifier.cpp:134-142, :222-235 */
orc_forest* orc_forest_read(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    orc_forest* F = (orc_forest*)calloc(1, sizeof(*F));
    if (!rd_i32(f, &F->T) || F->T <= 0) { fclose(f); free(F); return NULL; }
    F->trees = (orc_tree*)calloc(F->T, sizeof(orc_tree));
    for (int t = 0; t < F->T; t++) {
        orc_tree* tr = &F->trees[t];
        int n;
        rd_i32(f, &n); tr->n = n;
        tr->feat = (int*)malloc(4 * (size_t)n); if (fread(tr->feat, 4, n, f) != (size_t)n) goto bad;
        rd_i32(f, &n); if (n != tr->n) goto bad;
        tr->thr = (float*)malloc(4 * (size_t)n); if (fread(tr->thr, 4, n, f) != (size_t)n) goto bad;
        rd_i32(f, &n); if (n != tr->n) goto bad;
        tr->left = (int*)malloc(4 * (size_t)n); if (fread(tr->left, 4, n, f) != (size_t)n) goto bad;
        rd_i32(f, &n); if (n != tr->n) goto bad;
        for (int i = 0; i < n; i++) { /* plain histograms: skipped (empty in multi-label forests) */
            int c; rd_i32(f, &c);
            if (c > 0) fseek(f, 4L * c, SEEK_CUR);
        }
        rd_i32(f, &n); if (n != tr->n) goto bad;
        for (int i = 0; i < n; i++) {
            int L; rd_i32(f, &L);
            if (L > 0 && !F->C) {
                long here = ftell(f);
                F->L = L; F->C = (int*)malloc(4 * (size_t)L); F->sumC = 0;
                for (int l = 0; l < L; l++) { int c; rd_i32(f, &c); F->C[l] = c; F->sumC += c; fseek(f, 4L * c, SEEK_CUR); }
                fseek(f, here, SEEK_SET);
            }
            if (L > 0 && !tr->hist) tr->hist = (float*)calloc((size_t)tr->n * F->sumC, 4);
            float* h = tr->hist ? tr->hist + (size_t)i * F->sumC : NULL;
            for (int l = 0; l < L; l++) {
                int c; rd_i32(f, &c);
                if (l >= F->L || c != F->C[l]) goto bad;
                if (fread(h, 4, c, f) != (size_t)c) goto bad;
                h += c;
            }
        }
    }
    fclose(f);
    return F;
bad:
    fclose(f);
    orc_forest_free(F);
    return NULL;
}
void orc_forest_free(orc_forest* F) {
    if (!F) return;
    for (int t = 0; t < F->T; t++) { free(F->trees[t].feat); free(F->trees[t].thr); free(F->trees[t].left); free(F->trees[t].hist); }
    free(F->trees); free(F->C); free(F);
}
int orc_forest_trees(const orc_forest* F) { return F->T; }
int orc_forest_nodes(const orc_forest* F, int t) { return F->trees[t].n; }
int orc_forest_layers(const orc_forest* F) { return F->L; }
int orc_forest_classes(const orc_forest* F, int l) { return F->C[l]; }
/* classifier.cpp:97-117 */
static inline int find_leaf(const orc_tree* tr, const float* x) {
    int node = 0;
    while (tr->left[node] != 0) node = x[tr->feat[node]] < tr->thr[node] ? tr->left[node] : tr->left[node] + 1;
    return node;
}
/* classifier.cpp:187-208: tree 0's histogram, then trees 1..T-1 added in order */
void orc_forest_predict(const orc_forest* F, const float* feats, int n, int D, int* leaf_ids, float* logpost) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float* x = feats + (size_t)i * D;
        for (int t = 0; t < F->T; t++) {
            int leaf = find_leaf(&F->trees[t], x);
            if (leaf_ids) leaf_ids[(size_t)t * n + i] = leaf;
            if (logpost) {
                const float* h = F->trees[t].hist + (size_t)leaf * F->sumC;
                float* o = logpost + (size_t)i * F->sumC;
                if (t == 0) for (int c = 0; c < F->sumC; c++) o[c] = h[c];
                else for (int c = 0; c < F->sumC; c++) o[c] += h[c];
            }
        }
    }
}

/* src/segmenter.cpp:349-434 (fill = 0) and src/test_multi.cpp:166-199 (fill = -1000) */
void orc_segment_frame(const orc_fe_config* cfg, const orc_forest* F, int stride, const uint8_t* rgb,
                       const uint16_t* depth, int W, int H, const float* Kinv, const float* R, const float* t,
                       float dmin, float dmax, float fill, float* posteriors) {
    const int D = orc_feature_length(cfg);
    const int lw = W / stride, lh = H / stride;
    const size_t cap = (size_t)((W + stride - 1) / stride) * ((H + stride - 1) / stride);
    float* feats = (float*)malloc(cap * D * sizeof(float));
    int* xs = (int*)malloc(cap * sizeof(int));
    int* ys = (int*)malloc(cap * sizeof(int));
    int n = orc_extract(cfg, stride, rgb, depth, W, H, Kinv, R, t, dmin, dmax, ORC_NO_LABEL, NULL, 0, feats, xs, ys, NULL);
    float* post = (float*)malloc((size_t)n * F->sumC * sizeof(float));
    orc_forest_predict(F, feats, n, D, NULL, post);
    size_t off = 0;
    int coff = 0;
    for (int l = 0; l < F->L; l++) {
        const int C = F->C[l];
        float* low = (float*)malloc((size_t)lw * lh * C * sizeof(float));
        for (size_t k = 0; k < (size_t)lw * lh * C; k++) low[k] = fill;
        for (int j = 0; j < n; j++) { /* :366-376, p = row(y/stride) + C*x/stride */
            float* p = low + (size_t)(ys[j] / stride) * lw * C + (size_t)(C * xs[j] / stride);
            for (int c = 0; c < C; c++) p[c] = post[(size_t)j * F->sumC + coff + c];
        }
        orc_resize_linear_f32(low, lw, lh, C, posteriors + off, W, H); /* :380-382, then flattened :413-431 */
        free(low);
        off += (size_t)W * H * C;
        coff += C;
    }
    free(feats); free(xs); free(ys); free(post);
}

/* =============================================================================================
 * Permutohedral lattice: third-party/densecrf/src/permutohedral.cpp (the SSE build: x86-64 defines
 * __SSE__, so :140-321 init and :529-589 sseCompute are what the reference runs; :476-527 for M<=2)
 * ============================================================================================= */
struct orc_lattice {
    int N, d, V;
    int* offset;  /* [(N+16)*(d+1)] */
    float* bary;  /* [(N+16)*(d+1)] */
    int* nb;      /* [(d+1)][V][2] */
};
typedef struct {
    int ks;
    size_t cap, filled;
    short* keys;
    int* table;
} htab;
static size_t h_hash(const htab* h, const short* k) { /* :80-87 */
    size_t r = 0;
    for (int i = 0; i < h->ks; i++) { r += (size_t)(long)k[i]; r *= 1664525; }
    return r;
}
static void h_grow(htab* h) { /* :59-79 */
    size_t old_cap = h->cap;
    h->cap *= 2;
    h->keys = (short*)realloc(h->keys, sizeof(short) * (old_cap + 10) * h->ks);
    int* old = h->table;
    h->table = (int*)malloc(sizeof(int) * h->cap);
    for (size_t i = 0; i < h->cap; i++) h->table[i] = -1;
    for (size_t i = 0; i < old_cap; i++)
        if (old[i] >= 0) {
            int e = old[i];
            size_t p = h_hash(h, h->keys + (size_t)e * h->ks) % h->cap;
            while (h->table[p] >= 0) p = p < h->cap - 1 ? p + 1 : 0;
            h->table[p] = e;
        }
    free(old);
}
static int h_find(htab* h, const short* k, int create) { /* :98-127 */
    if (2 * h->filled >= h->cap) h_grow(h);
    size_t p = h_hash(h, k) % h->cap;
    for (;;) {
        int e = h->table[p];
        if (e == -1) {
            if (!create) return -1;
            memcpy(h->keys + h->filled * h->ks, k, sizeof(short) * h->ks);
            h->table[p] = (int)h->filled;
            return (int)h->filled++;
        }
        if (memcmp(h->keys + (size_t)e * h->ks, k, sizeof(short) * h->ks) == 0) return e;
        if (++p == h->cap) p = 0;
    }
}
orc_lattice* orc_lattice_init(const float* feature, int d, int N) {
    orc_lattice* L = (orc_lattice*)calloc(1, sizeof(*L));
    L->N = N; L->d = d;
    const int d1 = d + 1;
    L->offset = (int*)calloc((size_t)(N + 16) * d1, sizeof(int));
    L->bary = (float*)calloc((size_t)(N + 16) * d1, sizeof(float));
    htab h;
    h.ks = d; h.filled = 0; h.cap = 2 * (size_t)(N > 0 ? N : 1);
    h.keys = (short*)malloc(sizeof(short) * (h.cap / 2 + 10) * d);
    h.table = (int*)malloc(sizeof(int) * h.cap);
    for (size_t i = 0; i < h.cap; i++) h.table[i] = -1;

    const float invdplus1 = 1.0f / (float)d1, dplus1 = (float)d1;
    float scale_factor[32], elevated[33], rem0[33], rank[33], bary[34];
    short canonical[33 * 33], key[33];
    for (int i = 0; i <= d; i++) { /* :171-177 */
        for (int j = 0; j <= d - i; j++) canonical[i * d1 + j] = (short)i;
        for (int j = d - i + 1; j <= d; j++) canonical[i * d1 + j] = (short)(i - d1);
    }
    const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * d1);                                           /* :180 */
    for (int i = 0; i < d; i++) scale_factor[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev); /* :183 */

    /* the SSE loop handles 4 points per step and pads the last block with zero features (:195-198);
     * the padded points still insert their vertices, so they are walked here too. */
    const int Npad = (N + 3) & ~3;
    for (int p = 0; p < Npad; p++) {
        float sm = 0.f; /* :203-209 */
        for (int j = d; j > 0; j--) {
            float fv = p < N ? feature[(size_t)p * d + (j - 1)] : 0.0f;
            float cf = fv * scale_factor[j - 1];
            float jc = (float)j * cf;
            elevated[j] = sm - jc;
            sm = sm + cf;
        }
        elevated[0] = sm;
        float sum = 0.f; /* :212-222, cvtps_epi32 = round-half-even */
        for (int i = 0; i <= d; i++) {
            float v = invdplus1 * elevated[i];
            v = rintf(v);
            rem0[i] = v * dplus1;
            sum = sum + v;
        }
        for (int i = 0; i <= d; i++) rank[i] = 0.f; /* :225-235 */
        for (int i = 0; i < d; i++) {
            float di = elevated[i] - rem0[i];
            for (int j = i + 1; j <= d; j++) {
                float dj = elevated[j] - rem0[j];
                float c = di < dj ? 1.f : 0.f;
                rank[i] += c;
                rank[j] += 1.f - c;
            }
        }
        for (int i = 0; i <= d; i++) { /* :238-244 */
            rank[i] += sum;
            float add = rank[i] < 0.f ? dplus1 : 0.f, sub = rank[i] >= dplus1 ? dplus1 : 0.f;
            rank[i] += add - sub;
            rem0[i] += add - sub;
        }
        for (int i = 0; i < d + 2; i++) bary[i] = 0.f; /* :247-260 */
        for (int i = 0; i <= d; i++) {
            float v = (elevated[i] - rem0[i]) * invdplus1;
            int q = d - (int)rank[i];
            bary[q] += v;
            bary[q + 1] -= v;
        }
        bary[0] += 1.f + bary[d + 1]; /* :265 */
        for (int rem = 0; rem <= d; rem++) { /* :270-277 */
            for (int i = 0; i < d; i++) key[i] = (short)(rem0[i] + (float)canonical[rem * d1 + (int)rank[i]]);
            L->offset[(size_t)p * d1 + rem] = h_find(&h, key, 1);
            L->bary[(size_t)p * d1 + rem] = bary[rem];
        }
    }
    L->V = (int)h.filled;
    L->nb = (int*)malloc(sizeof(int) * 2 * (size_t)d1 * (L->V > 0 ? L->V : 1));
    short n1[33], n2[33];
    for (int j = 0; j <= d; j++) /* :303-318 */
        for (int i = 0; i < L->V; i++) {
            const short* k = h.keys + (size_t)i * d;
            for (int q = 0; q < d; q++) { n1[q] = (short)(k[q] - 1); n2[q] = (short)(k[q] + 1); }
            if (j < d) { n1[j] = (short)(k[j] + d); n2[j] = (short)(k[j] - d); }
            L->nb[((size_t)j * L->V + i) * 2] = h_find(&h, n1, 0);
            L->nb[((size_t)j * L->V + i) * 2 + 1] = h_find(&h, n2, 0);
        }
    free(h.keys); free(h.table);
    return L;
}
void orc_lattice_free(orc_lattice* L) { if (L) { free(L->offset); free(L->bary); free(L->nb); free(L); } }
int orc_lattice_vertices(const orc_lattice* L) { return L->V; }
void orc_lattice_get(const orc_lattice* L, int* offsets, float* bary) {
    memcpy(offsets, L->offset, sizeof(int) * (size_t)L->N * (L->d + 1));
    memcpy(bary, L->bary, sizeof(float) * (size_t)L->N * (L->d + 1));
}
/* :596-604 dispatch; :476-527 (rows<=2) and :529-589 (SSE, rows padded to a multiple of 4) */
void orc_lattice_compute(const orc_lattice* L, const float* in, int M, float* out) {
    const int d1 = L->d + 1, V = L->V, N = L->N;
    const int seq = M <= 2;
    const int vs = seq ? M : ((M - 1) / 4 + 1) * 4;
    float* val = (float*)calloc((size_t)(V + 2) * vs, sizeof(float));
    float* nval = (float*)calloc((size_t)(V + 2) * vs, sizeof(float));
    for (int i = 0; i < N; i++) /* splat */
        for (int j = 0; j < d1; j++) {
            size_t o = (size_t)(L->offset[(size_t)i * d1 + j] + 1) * vs;
            float w = L->bary[(size_t)i * d1 + j];
            for (int k = 0; k < M; k++) { float p = w * in[(size_t)i * M + k]; val[o + k] += p; }
        }
    for (int j = 0; j < d1; j++) { /* blur */
        for (int i = 0; i < V; i++) {
            const float* o = val + (size_t)(i + 1) * vs;
            float* nw = nval + (size_t)(i + 1) * vs;
            const float* a = val + (size_t)(L->nb[((size_t)j * V + i) * 2] + 1) * vs;
            const float* b = val + (size_t)(L->nb[((size_t)j * V + i) * 2 + 1] + 1) * vs;
            if (seq) for (int k = 0; k < vs; k++) { float s = a[k] + b[k]; nw[k] = (float)((double)o[k] + 0.5 * (double)s); }
            else for (int k = 0; k < vs; k++) { float s = a[k] + b[k]; float hs = 0.5f * s; nw[k] = o[k] + hs; }
        }
        float* tmp = val; val = nval; nval = tmp;
    }
    const float alpha = 1.0f / (1 + powf(2, -(float)L->d));
    for (int i = 0; i < N; i++) { /* slice */
        float acc[64];
        for (int k = 0; k < M; k++) acc[k] = 0.f;
        for (int j = 0; j < d1; j++) {
            size_t o = (size_t)(L->offset[(size_t)i * d1 + j] + 1) * vs;
            float w = L->bary[(size_t)i * d1 + j];
            if (seq) for (int k = 0; k < M; k++) { float p = w * val[o + k]; p = p * alpha; acc[k] += p; }
            else { float wa = w * alpha; for (int k = 0; k < M; k++) { float p = wa * val[o + k]; acc[k] += p; } }
        }
        for (int k = 0; k < M; k++) out[(size_t)i * M + k] = acc[k];
    }
    free(val); free(nval);
}

/* =============================================================================================
 * Mean field: densecrf.cpp:98-131, pairwise.cpp:40-80,173-178, labelcompatibility.cpp:46-48
 * ============================================================================================= */
static void exp_and_normalize(float* out, const float* in, int M, int N) { /* densecrf.cpp:98-106 */
    for (int i = 0; i < N; i++) {
        const float* b = in + (size_t)i * M;
        float* o = out + (size_t)i * M;
        float mx = b[0];
        for (int k = 1; k < M; k++) if (b[k] > mx) mx = b[k];
        float s = 0.f;
        for (int k = 0; k < M; k++) { o[k] = expf(b[k] - mx); s += o[k]; }
        for (int k = 0; k < M; k++) o[k] = o[k] / s;
    }
}
/* norm_type: pairwise.h NormalizationType (0 NO_NORMALIZATION, 1 NORMALIZE_BEFORE, 2 NORMALIZE_AFTER, 3 NORMALIZE_SYMMETRIC) */
void orc_crf_inference_ex(int N, int M, const float* unary, const orc_pairwise* kernels, int K, int iters, int norm_type,
                          float* Q) {
    orc_lattice** lat = (orc_lattice**)malloc(sizeof(void*) * (K > 0 ? K : 1));
    float** norm = (float**)malloc(sizeof(void*) * (K > 0 ? K : 1));
    float* ones = (float*)malloc(sizeof(float) * N);
    for (int i = 0; i < N; i++) ones[i] = 1.f;
    const int pre = norm_type == 3 || norm_type == 1;  /* pairwise.cpp:65: SYMMETRIC || (BEFORE && !transpose) */
    const int post = norm_type == 3 || norm_type == 2; /* pairwise.cpp:78: SYMMETRIC || (AFTER && !transpose) */
    for (int k = 0; k < K; k++) { /* pairwise.cpp:40-62 */
        lat[k] = orc_lattice_init(kernels[k].feats, kernels[k].d, N);
        norm[k] = (float*)malloc(sizeof(float) * N);
        orc_lattice_compute(lat[k], ones, 1, norm[k]);
        if (norm_type == 0) { /* :46-52: every point gets N / sum(norm); NO filter-time scaling (:65,:78) */
            float mean_norm = 0;
            for (int i = 0; i < N; i++) mean_norm += norm[k][i];
            mean_norm = N / mean_norm;
            for (int i = 0; i < N; i++) norm[k][i] = mean_norm;
        } else if (norm_type == 3) {
            for (int i = 0; i < N; i++) norm[k][i] = (float)(1.0 / sqrt((double)norm[k][i] + 1e-20));
        } else {
            for (int i = 0; i < N; i++) norm[k][i] = (float)(1.0 / ((double)norm[k][i] + 1e-20));
        }
    }
    float* tmp1 = (float*)malloc(sizeof(float) * (size_t)M * N);
    float* tmp2 = (float*)malloc(sizeof(float) * (size_t)M * N);
    for (size_t i = 0; i < (size_t)M * N; i++) tmp1[i] = -unary[i];
    exp_and_normalize(Q, tmp1, M, N); /* densecrf.cpp:120 */
    for (int it = 0; it < iters; it++) {
        for (size_t i = 0; i < (size_t)M * N; i++) tmp1[i] = -unary[i];
        for (int k = 0; k < K; k++) {
            for (int i = 0; i < N; i++) /* pairwise.cpp:65-68 out = in*norm, or out = in */
                for (int c = 0; c < M; c++) tmp2[(size_t)i * M + c] = pre ? Q[(size_t)i * M + c] * norm[k][i] : Q[(size_t)i * M + c];
            orc_lattice_compute(lat[k], tmp2, M, tmp2);
            for (int i = 0; i < N; i++)
                for (int c = 0; c < M; c++) {
                    float v = tmp2[(size_t)i * M + c];
                    if (post) v = v * norm[k][i];       /* pairwise.cpp:78-79 */
                    v = -kernels[k].potts_w * v;        /* labelcompatibility.cpp:46-48 */
                    tmp1[(size_t)i * M + c] -= v;       /* densecrf.cpp:126 */
                }
        }
        exp_and_normalize(Q, tmp1, M, N);
    }
    for (int k = 0; k < K; k++) { orc_lattice_free(lat[k]); free(norm[k]); }
    free(lat); free(norm); free(ones); free(tmp1); free(tmp2);
}
void orc_crf_inference(int N, int M, const float* unary, const orc_pairwise* kernels, int K, int iters, float* Q) {
    orc_crf_inference_ex(N, M, unary, kernels, K, iters, 3, Q);
}
/* The projector of src/segmenter.cpp:581 (fps_mapper's pinhole projector, un-vendored) as THIS project defines it: points
 * of the cloud (map frame) -> camera frame with the key-frame pose (R, t: camera -> map), pinhole projection with K
 * (row-major 3x3: fx 0 cx / 0 fy cy / 0 0 1), nearest pixel, z inside [zmin, zmax]; per pixel the index of the NEAREST point
 * wins (z-buffer; equal z: the lower index), -1 where nothing projects.  All arithmetic in float, every operation rounded
 * separately, in this order. */
void orc_project_zbuffer(const float* xyz, int N, const float* K, const float* R, const float* t, int W, int H, float zmin,
                         float zmax, int* index_image) {
    const float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
    float* zbuf = (float*)malloc(sizeof(float) * (size_t)W * H);
    for (size_t i = 0; i < (size_t)W * H; i++) { index_image[i] = -1; zbuf[i] = 0.f; }
    for (int i = 0; i < N; i++) {
        const float dx = xyz[3 * (size_t)i] - t[0], dy = xyz[3 * (size_t)i + 1] - t[1], dz = xyz[3 * (size_t)i + 2] - t[2];
        /* p_cam = R^T (p - t): column k of R dotted with d, ((a + b) + c) */
        float s;
        s = R[0] * dx; s = s + R[3] * dy; const float x = s + R[6] * dz;
        s = R[1] * dx; s = s + R[4] * dy; const float y = s + R[7] * dz;
        s = R[2] * dx; s = s + R[5] * dy; const float z = s + R[8] * dz;
        if (!(z >= zmin && z <= zmax)) continue;
        const float iz = 1.0f / z;
        float u = fx * x; u = u * iz; u = u + cx;
        float v = fy * y; v = v * iz; v = v + cy;
        const float fu = floorf(u + 0.5f), fv = floorf(v + 0.5f);
        if (!(fu >= 0.f && fu < (float)W && fv >= 0.f && fv < (float)H)) continue;
        const size_t pix = (size_t)(int)fv * W + (int)fu;
        if (index_image[pix] < 0 || z < zbuf[pix]) { index_image[pix] = i; zbuf[pix] = z; }
    }
    free(zbuf);
}
void orc_unary_accumulate(const int* index_image, int npix, const float* posterior, int C, float* unary) {
    for (int p = 0; p < npix; p++) { /* segmenter.cpp:599-616 */
        int idx = index_image[p];
        if (idx >= 0)
            for (int c = 0; c < C; c++) unary[(size_t)idx * C + c] += posterior[(size_t)p * C + c];
    }
}
void orc_gated_argmax(const float* Q, int M, int N, int unknown_label, uint8_t* labels) {
    for (int i = 0; i < N; i++) { /* segmenter.cpp:645-657 */
        unsigned mx = (unsigned)unknown_label;
        float mv = (float)(2.0 / M);
        for (int c = 0; c < M; c++) {
            float cur = Q[(size_t)i * M + c];
            if (cur > mv) { mv = cur; mx = (unsigned)c; }
        }
        labels[i] = (uint8_t)mx;
    }
}
void orc_features_gaussian2d(int W, int H, float sx, float sy, float* f) { /* densecrf.cpp:61-69 */
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            f[2 * ((size_t)j * W + i)] = i / sx;
            f[2 * ((size_t)j * W + i) + 1] = j / sy;
        }
}
void orc_features_bilateral2d(int W, int H, float sx, float sy, float sr, float sg, float sb, const uint8_t* im,
                              float* f) { /* densecrf.cpp:70-81 */
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            size_t p = (size_t)j * W + i;
            f[5 * p] = i / sx;
            f[5 * p + 1] = j / sy;
            f[5 * p + 2] = im[3 * p] / sr;
            f[5 * p + 3] = im[3 * p + 1] / sg;
            f[5 * p + 4] = im[3 * p + 2] / sb;
        }
}
void orc_features_xyzrgb(int N, const float* xyz, const float* rgb, float wxyz, float wrgb, float* f) {
    for (int i = 0; i < N; i++) /* segmenter.cpp:629-637 */
        for (int k = 0; k < 3; k++) {
            f[6 * (size_t)i + k] = xyz[3 * (size_t)i + k] * wxyz;
            f[6 * (size_t)i + 3 + k] = rgb[3 * (size_t)i + k] * wrgb;
        }
}
