// GPU side of the reference's EVAL input transforms: one kernel per image does what
//   albumentations.Resize(S, S, interpolation=cv2.INTER_CUBIC) -> albumentations.Normalize(mean, std) -> ToTensorV2
// (configs/experiment/coop/clipseg.yaml:113-127, applied in src/data/core_datasets/image_text_mask_dataset.py:52-84) do on
// the host, and one kernel resizes the mask the way albumentations does (cv2.INTER_NEAREST).
//
// The resize is OpenCV's 8-bit fixed-point cubic (modules/imgproc/src/resize.cpp: HResizeCubic / VResizeCubic with
// FixedPtCast<int, uchar, 22>): per output column / row four int16 taps (scaled by 2048) with replicated borders,
// horizontal sums in int32, vertical sum rounded with (v + 2^21) >> 22 and saturated to [0, 255].  The tap tables come from
// the host (tunevlseg_b200/data/gpu_transforms.py restates cv::resize's float coefficient computation); all device
// arithmetic is integer, so the uint8 image is bit-exact against oracle/preprocess.py.  Normalize is the float32 pair
// (x - 255 mean) then * 1 / (255 std), two roundings, no fused multiply-add.
#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

// one thread per output pixel, all three channels: 4 x 4 x 3 byte gathers (the source image sits in L2: <= a few MB)
__global__ void __launch_bounds__(256)
preproc_image_u8_kernel(const uint8_t* __restrict__ img, int Hi, int Wi, long long ld, const int* __restrict__ xofs, const int* __restrict__ xcoef,
                        const int* __restrict__ yofs, const int* __restrict__ ycoef, float m0, float m1, float m2, float d0, float d1,
                        float d2, int Ho, int Wo, float* __restrict__ out, uint8_t* __restrict__ out_u8) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= Wo) return;
    const int sx = xofs[x], sy = yofs[y];
    int cx[4], cy[4], ix[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cx[k] = xcoef[4 * x + k];
        cy[k] = ycoef[4 * y + k];
        ix[k] = min(max(sx + k - 1, 0), Wi - 1) * 3;
    }
    long long v[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint8_t* row = img + static_cast<long long>(min(max(sy + r - 1, 0), Hi - 1)) * ld;
        int h[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) h[c] += static_cast<int>(row[ix[k] + c]) * cx[k];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] += static_cast<long long>(h[c]) * cy[r];
    }
    const float mean[3] = {m0, m1, m2}, den[3] = {d0, d1, d2};
    const long long plane = static_cast<long long>(Ho) * Wo, pix = static_cast<long long>(y) * Wo + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long q = (v[c] + (1LL << 21)) >> 22;
        const int u = static_cast<int>(q < 0 ? 0 : (q > 255 ? 255 : q));
        if (out_u8) out_u8[pix * 3 + c] = static_cast<uint8_t>(u);
        if (out) out[c * plane + pix] = __fmul_rn(__fsub_rn(static_cast<float>(u), mean[c]), den[c]);
    }
}

__global__ void __launch_bounds__(256)
resize_nearest_f32_kernel(const float* __restrict__ in, long long ld, const int* __restrict__ xofs, const int* __restrict__ yofs, int Ho, int Wo,
                          float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= Wo) return;
    out[static_cast<long long>(y) * Wo + x] = in[static_cast<long long>(yofs[y]) * ld + xofs[x]];
}

// ---------------------------------------------------------------------------------------------------------------------
// TRAIN-time augmentations (configs/experiment/coop/clipseg.yaml:84-103): albumentations.Affine = cv2.warpAffine(INTER_CUBIC,
// BORDER_REPLICATE) on the resized uint8 image, RandomBrightnessContrast = a 256-entry uint8 LUT, then Normalize + ToTensorV2.
// warpAffine is fixed point end to end (modules/imgproc/src/imgwarp.cpp): the host inverts the matrix in double and builds the
// integer walk tables OpenCV builds (adelta / bdelta per column, X0 / Y0 per row, 1/1024 pixel, rounding term included); the
// device reduces to 5 fractional bits, looks the 4 x 4 int16 weights up in the 32 x 32 table of initInterTab2D (uploaded by
// the host, scaled by 2^15) and evaluates (sum + 2^14) >> 15 with saturation - every step integer, so the bytes equal cv2's.
__global__ void __launch_bounds__(256)
warp_affine_u8_kernel(const uint8_t* __restrict__ img, int Hi, int Wi, long long ld, const int* __restrict__ adelta, const int* __restrict__ bdelta,
                      const int* __restrict__ x0, const int* __restrict__ y0, const short* __restrict__ tab, const uint8_t* __restrict__ lut,
                      float m0, float m1, float m2, float d0, float d1, float d2, int Ho, int Wo, float* __restrict__ out,
                      uint8_t* __restrict__ out_u8) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= Wo) return;
    const int X = (x0[y] + adelta[x]) >> 5, Y = (y0[y] + bdelta[x]) >> 5;          // AB_BITS - INTER_BITS
    const int sx = min(max(X >> 5, -32768), 32767) - 1, sy = min(max(Y >> 5, -32768), 32767) - 1;   // saturate_cast<short>, first tap
    const uint4* wp = reinterpret_cast<const uint4*>(tab + (((Y & 31) << 5) + (X & 31)) * 16);
    const uint4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
    const uint32_t wu[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    int ix[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ix[k] = min(max(sx + k, 0), Wi - 1) * 3;
    int v[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint8_t* row = img + static_cast<long long>(min(max(sy + r, 0), Hi - 1)) * ld;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t pair = wu[(4 * r + k) >> 1];
            const int w = static_cast<short>((k & 1) ? (pair >> 16) : (pair & 0xffffu));
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] += static_cast<int>(row[ix[k] + c]) * w;
        }
    }
    const float mean[3] = {m0, m1, m2}, den[3] = {d0, d1, d2};
    const long long plane = static_cast<long long>(Ho) * Wo, pix = static_cast<long long>(y) * Wo + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int q = (v[c] + (1 << 14)) >> 15;
        int u = q < 0 ? 0 : (q > 255 ? 255 : q);
        if (lut) u = lut[u];
        if (out_u8) out_u8[pix * 3 + c] = static_cast<uint8_t>(u);
        if (out) out[c * plane + pix] = __fmul_rn(__fsub_rn(static_cast<float>(u), mean[c]), den[c]);
    }
}

__global__ void __launch_bounds__(256)
warp_affine_nearest_f32_kernel(const float* __restrict__ in, int Hi, int Wi, long long ld, const int* __restrict__ adelta, const int* __restrict__ bdelta,
                               const int* __restrict__ x0, const int* __restrict__ y0, int Ho, int Wo, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= Wo) return;
    const int sx = min(max((x0[y] + adelta[x]) >> 10, 0), Wi - 1), sy = min(max((y0[y] + bdelta[x]) >> 10, 0), Hi - 1);
    out[static_cast<long long>(y) * Wo + x] = in[static_cast<long long>(sy) * ld + sx];
}

// RandomBrightnessContrast without an Affine in front of it: LUT + Normalize + ToTensorV2 on the (already resized) uint8 image
__global__ void __launch_bounds__(256)
lut_normalize_u8_kernel(const uint8_t* __restrict__ img, long long ld, const uint8_t* __restrict__ lut, float m0, float m1, float m2, float d0,
                        float d1, float d2, int H, int W, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float mean[3] = {m0, m1, m2}, den[3] = {d0, d1, d2};
    const long long plane = static_cast<long long>(H) * W, pix = static_cast<long long>(y) * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int u = img[static_cast<long long>(y) * ld + 3 * x + c];
        if (lut) u = lut[u];
        out[c * plane + pix] = __fmul_rn(__fsub_rn(static_cast<float>(u), mean[c]), den[c]);
    }
}

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int tvs_preproc_image_u8(const uint8_t* img, int32_t Hi, int32_t Wi, int64_t ld_bytes, const int32_t* xofs,
                                                                            const int32_t* xcoef, const int32_t* yofs, const int32_t* ycoef,
                                                                            const float* mean255, const float* inv_std255, int32_t Ho, int32_t Wo,
                                                                            float* out_chw, uint8_t* out_u8_hwc, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(img && xofs && xcoef && yofs && ycoef && (out_chw || out_u8_hwc), "tvs_preproc_image_u8: null pointer");
    TVS_REQUIRE(!out_chw || (mean255 && inv_std255), "tvs_preproc_image_u8: mean255 / inv_std255 (host pointers to 3 floats) are needed for the float output");
    TVS_REQUIRE(Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && ld_bytes >= 3LL * Wi, "tvs_preproc_image_u8: bad geometry %dx%d -> %dx%d (ld %lld)", Hi, Wi, Ho, Wo,
                (long long)ld_bytes);
    const float m[3] = {mean255 ? mean255[0] : 0.f, mean255 ? mean255[1] : 0.f, mean255 ? mean255[2] : 0.f};
    const float d[3] = {inv_std255 ? inv_std255[0] : 1.f, inv_std255 ? inv_std255[1] : 1.f, inv_std255 ? inv_std255[2] : 1.f};
    preproc_image_u8_kernel<<<dim3((Wo + 255) / 256, Ho), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, Hi, Wi, ld_bytes, xofs, xcoef, yofs, ycoef, m[0], m[1],
                                                                                                       m[2], d[0], d[1], d[2], Ho, Wo, out_chw, out_u8_hwc);
    return check_launch("preproc_image_u8_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_resize_nearest_f32(const float* in, int32_t Hi, int32_t Wi, int64_t ld, const int32_t* xofs,
                                                                              const int32_t* yofs, int32_t Ho, int32_t Wo, float* out, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(in && xofs && yofs && out, "tvs_resize_nearest_f32: null pointer");
    TVS_REQUIRE(Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && ld >= Wi, "tvs_resize_nearest_f32: bad geometry");
    resize_nearest_f32_kernel<<<dim3((Wo + 255) / 256, Ho), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, ld, xofs, yofs, Ho, Wo, out);
    return check_launch("resize_nearest_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_warp_affine_u8(const uint8_t* img, int32_t Hi, int32_t Wi, int64_t ld_bytes, const int32_t* adelta,
                                                                          const int32_t* bdelta, const int32_t* x0, const int32_t* y0, const int16_t* tab,
                                                                          const uint8_t* lut, const float* mean255, const float* inv_std255, int32_t Ho,
                                                                          int32_t Wo, float* out_chw, uint8_t* out_u8_hwc, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(img && adelta && bdelta && x0 && y0 && tab && (out_chw || out_u8_hwc), "tvs_warp_affine_u8: null pointer");
    TVS_REQUIRE(!out_chw || (mean255 && inv_std255), "tvs_warp_affine_u8: mean255 / inv_std255 (host pointers to 3 floats) are needed for the float output");
    TVS_REQUIRE(Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && ld_bytes >= 3LL * Wi, "tvs_warp_affine_u8: bad geometry %dx%d -> %dx%d (ld %lld)", Hi, Wi, Ho, Wo,
                (long long)ld_bytes);
    TVS_REQUIRE((reinterpret_cast<uintptr_t>(tab) & 15) == 0, "tvs_warp_affine_u8: the weight table must be 16-byte aligned");
    const float m[3] = {mean255 ? mean255[0] : 0.f, mean255 ? mean255[1] : 0.f, mean255 ? mean255[2] : 0.f};
    const float d[3] = {inv_std255 ? inv_std255[0] : 1.f, inv_std255 ? inv_std255[1] : 1.f, inv_std255 ? inv_std255[2] : 1.f};
    warp_affine_u8_kernel<<<dim3((Wo + 255) / 256, Ho), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, Hi, Wi, ld_bytes, adelta, bdelta, x0, y0, tab, lut, m[0], m[1],
                                                                                                     m[2], d[0], d[1], d[2], Ho, Wo, out_chw, out_u8_hwc);
    return check_launch("warp_affine_u8_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_warp_affine_nearest_f32(const float* in, int32_t Hi, int32_t Wi, int64_t ld, const int32_t* adelta,
                                                                                   const int32_t* bdelta, const int32_t* x0, const int32_t* y0, int32_t Ho,
                                                                                   int32_t Wo, float* out, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(in && adelta && bdelta && x0 && y0 && out, "tvs_warp_affine_nearest_f32: null pointer");
    TVS_REQUIRE(Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && ld >= Wi, "tvs_warp_affine_nearest_f32: bad geometry");
    warp_affine_nearest_f32_kernel<<<dim3((Wo + 255) / 256, Ho), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, Hi, Wi, ld, adelta, bdelta, x0, y0, Ho, Wo, out);
    return check_launch("warp_affine_nearest_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_lut_normalize_u8(const uint8_t* img, int32_t H, int32_t W, int64_t ld_bytes, const uint8_t* lut,
                                                                            const float* mean255, const float* inv_std255, float* out_chw, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(img && mean255 && inv_std255 && out_chw, "tvs_lut_normalize_u8: null pointer");
    TVS_REQUIRE(H > 0 && W > 0 && H <= 65535 && ld_bytes >= 3LL * W, "tvs_lut_normalize_u8: bad geometry");
    lut_normalize_u8_kernel<<<dim3((W + 255) / 256, H), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, ld_bytes, lut, mean255[0], mean255[1], mean255[2], inv_std255[0],
                                                                                                     inv_std255[1], inv_std255[2], H, W, out_chw);
    return check_launch("lut_normalize_u8_kernel");
}
