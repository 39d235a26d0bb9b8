// On-GPU input pipeline: what the reference does per dataset item on the host, in one
// pass over the uint8 frames a decoder / resize hands over.
//
//   to_tensor                              mono/datasets/mono_dataset.py:100        uint8 HWC -> fp32 CHW, value / 255
//   color_aug = transforms.ColorJitter     mono/datasets/mono_dataset.py:62-73,182-187,102
//       torchvision ColorJitter on PIL images: brightness / contrast / saturation / hue in a sampled order, every
//       operation on BYTES (Pillow: Image.blend with truncation, convert("L"), convert("HSV") and back)
//   erase mask                             mono/datasets/kitti_dataset.py:167-182   ones with erase_count zeroed boxes
//   horizontal flip                        mono/datasets/mono_dataset.py:141,203-207 (applied here to the resized frame)
//
// The byte arithmetic follows Pillow's C code operation by operation (float where it uses float, double where it uses
// double, explicit round-to-nearest intrinsics so that nothing is contracted into an fma): the outputs are BIT-EXACT
// against torchvision 0.26 / Pillow 12.2 (oracle/input_pipeline.py restates it, tests/test_input_pipeline.py pins the
// restatement to the libraries over all 2^24 colours).
//
// Two launches: input_stat_kernel (only when some image needs it) sums the grey values ImageEnhance.Contrast averages --
// of the image as it is when the contrast step is reached in that image's order, so the preceding steps are replayed per
// pixel -- with integer atomics (exact, order-independent); input_apply_kernel then reads each pixel's three bytes once
// and writes color, color_aug and the mask.  HBM-bound by design: 3 B read (+3 B for the statistics pass), 24 B (+12 B
// mask on frame 0) written per pixel and frame.
#include "tdl_common.cuh"
#include "tdl_internal.h"

namespace tdl {

namespace {

struct Rgb {
    int r, g, b;
};

TDL_DEV int gray_u8(const Rgb& c) { return (c.r * 19595 + c.g * 38470 + c.b * 7471 + 0x8000) >> 16; }

// Pillow Blend.c: (UINT8)((int)deg + alpha * ((int)img - (int)deg)), clipped first when alpha is outside [0, 1]
TDL_DEV int blend1(int deg, int img, float alpha, bool inside) {
    const float t = __fadd_rn((float)deg, __fmul_rn(alpha, (float)(img - deg)));
    if (inside) return (int)t & 255;                       // 0 <= t <= 255 by construction
    return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}
TDL_DEV Rgb blend3(const Rgb& deg, const Rgb& img, float alpha) {
    const bool inside = alpha >= 0.f && alpha <= 1.f;
    return Rgb{blend1(deg.r, img.r, alpha, inside), blend1(deg.g, img.g, alpha, inside), blend1(deg.b, img.b, alpha, inside)};
}

TDL_DEV int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Pillow Convert.c rgb2hsv_row / hsv2rgb_row around torchvision's wrapping byte addition on H
TDL_DEV Rgb hue_shift(const Rgb& c, int shift) {
    const int maxc = max(c.r, max(c.g, c.b)), minc = min(c.r, min(c.g, c.b));
    int uh = 0, us = 0;
    const int uv = maxc;
    if (minc != maxc) {
        const float cr = (float)(maxc - minc);
        const float s = __fdiv_rn(cr, (float)maxc);
        const float rc = __fdiv_rn((float)(maxc - c.r), cr), gc = __fdiv_rn((float)(maxc - c.g), cr),
                    bc = __fdiv_rn((float)(maxc - c.b), cr);
        float h;
        if (c.r == maxc) h = __fsub_rn(bc, gc);
        else if (c.g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
        else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
        h = (float)fmod(__dadd_rn(__ddiv_rn((double)h, 6.0), 1.0), 1.0);
        uh = clip8((int)__dmul_rn((double)h, 255.0));
        us = clip8((int)__dmul_rn((double)s, 255.0));
    }
    uh = (uh + shift) & 255;
    if (us == 0) return Rgb{uv, uv, uv};
    const double fs = __ddiv_rn((double)us, 255.0);
    const double fh = __ddiv_rn(__dmul_rn((double)uh, 6.0), 255.0);
    const double fi = floor(fh);
    const double f = __dsub_rn(fh, fi);
    const double v = (double)uv;
    const int p = clip8((int)round(__dmul_rn(v, __dsub_rn(1.0, fs))));
    const int q = clip8((int)round(__dmul_rn(v, __dsub_rn(1.0, __dmul_rn(fs, f)))));
    const int t = clip8((int)round(__dmul_rn(v, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f))))));
    switch ((int)fi % 6) {
        case 0: return Rgb{uv, t, p};
        case 1: return Rgb{q, uv, p};
        case 2: return Rgb{p, uv, t};
        case 3: return Rgb{p, q, uv};
        case 4: return Rgb{t, p, uv};
        default: return Rgb{uv, p, q};
    }
}

struct Jitter {                // the sampled ColorJitter parameters of one (image, frame)
    float bright, contrast, sat;
    int hue;                   // byte added to H
    int order[4];              // fn_idx: 0 brightness, 1 contrast, 2 saturation, 3 hue
};

TDL_DEV Jitter load_jitter(const InputDev& p, int b, int fr) {
    Jitter j;
    const float* q = p.jitter + ((size_t)b * p.nframes + fr) * 4;
    j.bright = __ldg(q);
    j.contrast = __ldg(q + 1);
    j.sat = __ldg(q + 2);
    j.hue = (int)__ldg(q + 3) & 255;
    const int* o = p.order + ((size_t)b * p.nframes + fr) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) j.order[k] = __ldg(o + k);
    return j;
}

// the colour operations of positions [0, stop) -- or all four -- of this image's order; `mean` = Contrast's degenerate grey
TDL_DEV Rgb apply_chain(Rgb c, const Jitter& j, int stop, int mean) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= stop) break;
        const int fn = j.order[k];
        if (fn == 0) c = blend3(Rgb{0, 0, 0}, c, j.bright);
        else if (fn == 1) c = blend3(Rgb{mean, mean, mean}, c, j.contrast);
        else if (fn == 2) {
            const int g = gray_u8(c);
            c = blend3(Rgb{g, g, g}, c, j.sat);
        } else if (fn == 3) c = hue_shift(c, j.hue);
    }
    return c;
}

TDL_DEV Rgb load_px(const InputDev& p, int fr, int b, int y, int x, bool flip) {
    const int sx = flip ? p.W - 1 - x : x;
    const unsigned char* q = p.frames[fr] + (((size_t)b * p.H + y) * p.W + sx) * 3;
    return Rgb{(int)q[0], (int)q[1], (int)q[2]};
}

// grey sum of (image, frame) after the operations that precede the contrast step: gsum[b][frame] (exact integers)
__global__ void __launch_bounds__(256) input_stat_kernel(const InputDev p) {
    const int b = blockIdx.z, fr = blockIdx.y;
    if (!p.do_aug[b]) return;
    const Jitter j = load_jitter(p, b, fr);
    int cpos = 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (j.order[k] == 1) cpos = min(cpos, k);
    if (cpos == 4) return;                                  // no contrast step in this order
    const int n = p.H * p.W;
    unsigned int acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / p.W, x = i - y * p.W;
        const Rgb c = apply_chain(load_px(p, fr, b, y, x, false), j, cpos, 0);     // (the sum does not depend on the flip)
        acc += (unsigned)gray_u8(c);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    __shared__ unsigned int s_acc[8];
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_acc[k];
        atomicAdd(p.gsum + (size_t)b * p.nframes + fr, t);
    }
}

__global__ void __launch_bounds__(256) input_apply_kernel(const InputDev p) {
    const int b = blockIdx.z, fr = blockIdx.y;
    const int n = p.H * p.W;
    const bool flip = p.do_flip && p.do_flip[b];
    const bool aug = p.jitter && p.do_aug[b];
    Jitter j = {};
    int mean = 0;
    if (aug) {
        j = load_jitter(p, b, fr);
        // int(sum / count + 0.5) in double precision (PIL.ImageStat + ImageEnhance.Contrast)
        mean = (int)__dadd_rn(__ddiv_rn((double)p.gsum[(size_t)b * p.nframes + fr], (double)n), 0.5);
    }
    const size_t plane = (size_t)n;
    float* co = p.color[fr] ? p.color[fr] + (size_t)b * 3 * plane : nullptr;
    float* ca = p.color_aug[fr] ? p.color_aug[fr] + (size_t)b * 3 * plane : nullptr;
    float* mk = (fr == 0 && p.mask) ? p.mask + (size_t)b * 3 * plane : nullptr;
    const int* holes = p.holes ? p.holes + (size_t)b * p.erase_count * 2 : nullptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / p.W, x = i - y * p.W;
        const Rgb c = load_px(p, fr, b, y, x, flip);
        const float r = __fdiv_rn((float)c.r, 255.f), g = __fdiv_rn((float)c.g, 255.f), bl = __fdiv_rn((float)c.b, 255.f);
        if (co) {
            __stcs(co + i, r);
            __stcs(co + plane + i, g);
            __stcs(co + 2 * plane + i, bl);
        }
        if (ca) {
            if (aug) {
                const Rgb a = apply_chain(c, j, 4, mean);
                __stcs(ca + i, __fdiv_rn((float)a.r, 255.f));
                __stcs(ca + plane + i, __fdiv_rn((float)a.g, 255.f));
                __stcs(ca + 2 * plane + i, __fdiv_rn((float)a.b, 255.f));
            } else {
                __stcs(ca + i, r);
                __stcs(ca + plane + i, g);
                __stcs(ca + 2 * plane + i, bl);
            }
        }
        if (mk) {
            float m = 1.f;
            for (int k = 0; k < p.erase_count; ++k) {
                const int row = __ldg(holes + 2 * k), col = __ldg(holes + 2 * k + 1);
                if (y >= row && y < row + p.erase_h && x >= col && x < col + p.erase_w) m = 0.f;
            }
            __stcs(mk + i, m);
            __stcs(mk + plane + i, m);
            __stcs(mk + 2 * plane + i, m);
        }
    }
}

}  // namespace

cudaError_t launch_input_stat(const InputDev& p, cudaStream_t st) {
    const int n = p.H * p.W;
    dim3 grid(min((n + 255) / 256, 148 * 2), p.nframes, p.B);
    input_stat_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_input_apply(const InputDev& p, cudaStream_t st) {
    const int n = p.H * p.W;
    dim3 grid(min((n + 255) / 256, 148 * 4), p.nframes, p.B);
    input_apply_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace tdl
