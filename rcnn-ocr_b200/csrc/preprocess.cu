// K7: the host->device input step of the reference, one launch per batch of line images.
//
// Replaces, per image, data/transforms.py:62-120 (ResizeAndPadA: aspect-preserving cv2.resize onto a white
// img_h x img_w canvas, INTER_AREA when shrinking / INTER_LINEAR when enlarging), data/transforms.py:179
// (A.Normalize(mean 0.5, std 0.5): (v - 127.5) * (1 / 127.5)), ToTensorV2 (HWC -> CHW) and the per-image
// `.to(device)` + torch.stack of inference.py:93-124,159-164.  The decoded images of a batch travel as ONE pinned
// uint8 buffer (whatever sizes they have) + a descriptor table; every output pixel is computed where it is needed.
//
// Arithmetic follows OpenCV's 8-bit resize so that a model sees the pixels the reference showed it:
//   INTER_LINEAR   fixed point: 11-bit coefficients (saturate_cast<short>(f * 2048)), horizontal pass into a 32-bit
//                  intermediate, vertical pass ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
//   INTER_AREA     integer scale factors: box sum * (1 / area) (2x2: (sum + 2) >> 2); otherwise the fractional-coverage
//                  weights of cv::resizeArea_, accumulated in float, rounded half to even
// Parity bar (tests/test_preprocess_gpu.py): every output within ONE 8-bit level (2/255 after Normalize) of
// cv2 + the reference's formulae, and at least 97 % of the values exactly equal -- OpenCV's SIMD paths differ from its own
// scalar code by an 8-bit level at exact .5 cases, so the bar is not bit-exactness.
// Bound: HBM (a few source bytes and 4 / 2 output bytes per pixel); at 32 x 128 images the launch is latency-bound.
#include <math.h>
#include "common.cuh"

namespace rcnn {
namespace {

struct ImgDesc {          // int64 x 6 per image (host-packed)
    long long offset;     // byte offset of the first pixel in the packed buffer
    long long h, w;       // source height / width
    long long pitch;      // bytes per source row
    long long channels;   // 1 (grey), 3, 4 (alpha dropped)
    long long bgr;        // 1: the channel order is BGR(A) (cv2.imread)
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int sat_short(float v) {
    int i = __float2int_rn(v);           // cvRound
    return i < -32768 ? -32768 : (i > 32767 ? 32767 : i);
}

struct Src {
    const unsigned char *base;
    int pitch, channels, bgr;
    __device__ __forceinline__ int at(int y, int x, int c) const {
        if (channels == 1) return base[(size_t)y * pitch + x];
        const int cc = bgr ? 2 - c : c;
        return base[(size_t)y * pitch + (size_t)x * channels + cc];
    }
};

// cv::resize INTER_LINEAR, 8-bit: coordinate map and 11-bit coefficients of one axis
__device__ __forceinline__ void linear_coef(int d, double scale, int ssize, int &s0, int &c0, int &c1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    s0 = s;
    c0 = sat_short((1.f - f) * 2048.f);
    c1 = sat_short(f * 2048.f);
}

template <typename OutT> __device__ __forceinline__ OutT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// grid: (output row, image); block: threads stride over the output columns
template <typename OutT>
__global__ void preprocess_lines_kernel(const unsigned char *__restrict__ pixels, const ImgDesc *__restrict__ desc,
                                        int img_h, int img_w, int align_h, int align_v, OutT *__restrict__ out) {
    const int n = blockIdx.y, oy = blockIdx.x;
    const ImgDesc d = desc[n];
    const int h = (int)d.h, w = (int)d.w;
    Src src{pixels + d.offset, (int)d.pitch, (int)d.channels, (int)d.bgr};
    // ResizeAndPadA.apply: scale, new size (python round = half to even), interpolation, placement
    const double sc = fmin((double)img_h / (double)max(h, 1), (double)img_w / (double)max(w, 1));
    const int new_w = max(1, (int)rint((double)w * sc)), new_h = max(1, (int)rint((double)h * sc));
    const bool area = new_h < h || new_w < w;
    int x0 = align_h == 0 ? 0 : (align_h == 2 ? img_w - new_w : (img_w - new_w) / 2);
    int y0 = align_v == 0 ? 0 : (align_v == 2 ? img_h - new_h : (img_h - new_h) / 2);
    x0 = max(0, min(x0, img_w - new_w));
    y0 = max(0, min(y0, img_h - new_h));
    const float kInv = 0.007843138f;      // np.reciprocal(np.float32(127.5)) as albumentations forms it
    OutT *orow = out + ((size_t)n * 3 * img_h + oy) * img_w;
    const size_t plane = (size_t)img_h * img_w;
    const int dy = oy - y0;
    const bool row_in = dy >= 0 && dy < new_h;
    const double scale_x = (double)w / new_w, scale_y = (double)h / new_h;   // cv::resize: inv_scale = dsize / ssize

    // vertical taps of this output row
    int sy0 = 0, by0 = 0, by1 = 0;
    int ay1 = 0, ay2 = 0;
    float wy_first = 0.f, wy_mid = 0.f, wy_last = 0.f;
    const int iscale_x = (int)rint(scale_x), iscale_y = (int)rint(scale_y);
    const bool area_fast = area && fabs(scale_x - iscale_x) < 2.220446049250313e-16 && fabs(scale_y - iscale_y) < 2.220446049250313e-16;
    bool have_first_y = false, have_last_y = false;
    if (row_in) {
        if (!area) {
            linear_coef(dy, scale_y, h, sy0, by0, by1);
        } else if (!area_fast) {
            // cv::computeResizeAreaTab for this dy
            const double fsy1 = dy * scale_y, fsy2 = fsy1 + scale_y;
            const double cell = fmin(scale_y, (double)h - fsy1);
            int s1 = (int)ceil(fsy1), s2 = (int)floor(fsy2);
            s2 = min(s2, h - 1);
            s1 = min(s1, s2);
            ay1 = s1; ay2 = s2;
            have_first_y = (s1 - fsy1) > 1e-3;
            wy_first = (float)((s1 - fsy1) / cell);
            wy_mid = (float)(1.0 / cell);
            have_last_y = (fsy2 - s2) > 1e-3;
            wy_last = (float)(fmin(fmin(fsy2 - s2, 1.0), cell) / cell);
        }
    }

    for (int ox = threadIdx.x; ox < img_w; ox += blockDim.x) {
        const int dx = ox - x0;
        float v[3] = {255.f, 255.f, 255.f};         // the white canvas
        if (row_in && dx >= 0 && dx < new_w) {
            if (!area) {
                int sx0, ax0, ax1;
                linear_coef(dx, scale_x, w, sx0, ax0, ax1);
                const int sx1 = min(sx0 + 1, w - 1), sy1 = min(sy0 + 1, h - 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int r0 = src.at(sy0, sx0, c) * ax0 + src.at(sy0, sx1, c) * ax1;
                    const int r1 = src.at(sy1, sx0, c) * ax0 + src.at(sy1, sx1, c) * ax1;
                    const int q = (((by0 * (r0 >> 4)) >> 16) + ((by1 * (r1 >> 4)) >> 16) + 2) >> 2;
                    v[c] = (float)clampi(q, 0, 255);
                }
            } else if (area_fast) {
                const int bx = dx * iscale_x, by = dy * iscale_y;
                const int area_n = iscale_x * iscale_y;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int sum = 0;
                    for (int yy = 0; yy < iscale_y; ++yy)
                        for (int xx = 0; xx < iscale_x; ++xx) sum += src.at(min(by + yy, h - 1), min(bx + xx, w - 1), c);
                    int q;
                    if (iscale_x == 2 && iscale_y == 2) q = (sum + 2) >> 2;
                    else q = __float2int_rn((float)sum * (1.f / (float)area_n));
                    v[c] = (float)clampi(q, 0, 255);
                }
            } else {
                const double fsx1 = dx * scale_x, fsx2 = fsx1 + scale_x;
                const double cell = fmin(scale_x, (double)w - fsx1);
                int s1 = (int)ceil(fsx1), s2 = (int)floor(fsx2);
                s2 = min(s2, w - 1);
                s1 = min(s1, s2);
                const bool first_x = (s1 - fsx1) > 1e-3, last_x = (fsx2 - s2) > 1e-3;
                const float wx_first = (float)((s1 - fsx1) / cell), wx_mid = (float)(1.0 / cell);
                const float wx_last = (float)(fmin(fmin(fsx2 - s2, 1.0), cell) / cell);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float acc = 0.f;
                    auto hrow = [&](int sy) {      // buf[dx] of resizeArea_: horizontal weighted sum of one source row
                        float b = 0.f;      // (separate multiply and add roundings, as the scalar reference code)
                        if (first_x) b = __fadd_rn(b, __fmul_rn((float)src.at(sy, s1 - 1, c), wx_first));
                        for (int sx = s1; sx < s2; ++sx) b = __fadd_rn(b, __fmul_rn((float)src.at(sy, sx, c), wx_mid));
                        if (last_x) b = __fadd_rn(b, __fmul_rn((float)src.at(sy, s2, c), wx_last));
                        return b;
                    };
                    bool started = false;
                    if (have_first_y) { acc = __fmul_rn(wy_first, hrow(ay1 - 1)); started = true; }
                    for (int sy = ay1; sy < ay2; ++sy) {
                        const float t = __fmul_rn(wy_mid, hrow(sy));
                        acc = started ? __fadd_rn(acc, t) : t;
                        started = true;
                    }
                    if (have_last_y) {
                        const float t = __fmul_rn(wy_last, hrow(ay2));
                        acc = started ? __fadd_rn(acc, t) : t;
                    }
                    v[c] = (float)clampi(__float2int_rn(acc), 0, 255);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) orow[(size_t)c * plane + ox] = to_out<OutT>(__fmul_rn(v[c] - 127.5f, kInv));
    }
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_preprocess_lines(const void *pixels, const int64_t *desc, int N, int img_h, int img_w, int align_h,
                                     int align_v, void *out, int out_dtype, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(N >= 0 && img_h > 0 && img_w > 0, "preprocess_lines: bad shape N=%d img %dx%d", N, img_h, img_w);
    RCNN_CHECK_ARG(align_h >= 0 && align_h <= 2 && align_v >= 0 && align_v <= 2, "preprocess_lines: alignment codes are 0 (left/top), 1 (center), 2 (right/bottom)");
    RCNN_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "preprocess_lines: out_dtype %d (0 = f32, 1 = bf16)", out_dtype);
    RCNN_CHECK_ARG(img_h <= 65535, "preprocess_lines: img_h %d too large", img_h);
    if (N == 0) return RCNN_OK;
    RCNN_CHECK_ARG(pixels && desc && out, "preprocess_lines: null pointer");
    static_assert(sizeof(ImgDesc) == 6 * sizeof(int64_t), "descriptor layout");
    const int threads = img_w >= 256 ? 256 : (img_w >= 128 ? 128 : 64);
    dim3 grid((unsigned)img_h, (unsigned)N);
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == 0)
        preprocess_lines_kernel<float><<<grid, threads, 0, s>>>((const unsigned char *)pixels, (const ImgDesc *)desc, img_h, img_w,
                                                                 align_h, align_v, (float *)out);
    else
        preprocess_lines_kernel<__nv_bfloat16><<<grid, threads, 0, s>>>((const unsigned char *)pixels, (const ImgDesc *)desc, img_h,
                                                                         img_w, align_h, align_v, (__nv_bfloat16 *)out);
    RCNN_LAUNCH_CHECK("preprocess_lines_kernel");
    return RCNN_OK;
}
