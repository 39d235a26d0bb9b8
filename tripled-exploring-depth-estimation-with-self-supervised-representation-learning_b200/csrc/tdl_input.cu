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

constexpr int kMaxBoxes = 64;      // erase boxes per image held in shared memory (cfg_kitti_tripleD: 16)

struct Rgb {
    int r, g, b;
};

TDL_DEV int gray_u8(const Rgb& c) { return (c.r * 19595 + c.g * 38470 + c.b * 7471 + 0x8000) >> 16; }

// Pillow Blend.c: (UINT8)((int)deg + alpha * ((int)img - (int)deg)), clipped to [0, 255] first when alpha is outside
// [0, 1] -- inside that range the value already lies in [0, 255], so one clamped truncation serves both branches
TDL_DEV int blend1(int deg, int img, float alpha) {
    const float t = __fadd_rn((float)deg, __fmul_rn(alpha, (float)(img - deg)));
    return __float2int_rz(fminf(fmaxf(t, 0.f), 255.f));
}
TDL_DEV Rgb blend3(const Rgb& deg, const Rgb& img, float alpha) {
    return Rgb{blend1(deg.r, img.r, alpha), blend1(deg.g, img.g, alpha), blend1(deg.b, img.b, alpha)};
}

TDL_DEV int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Per-CTA tables of the exact quotients the byte arithmetic needs (filled once per CTA with IEEE divisions): a correctly
// rounded division costs ~10 (float) / ~40 (double) instructions, a shared-memory load one.
struct Tables {
    float q255f[256];     // (float)x / 255.f                  transforms.ToTensor
    float rcpf[256];      // RN(1 / (float)x)                   seed of the correctly rounded float quotients below
    double q255d[256];    // (double)x / 255.0                  hsv2rgb: fs
    double frac[256];     // fh - floor(fh), fh = x * 6.0 / 255.0
    unsigned char sect[256];   // floor(fh) % 6
};

TDL_DEV void fill_tables(Tables& t, int tid) {
    if (tid < 256) {
        t.q255f[tid] = __fdiv_rn((float)tid, 255.f);
        t.rcpf[tid] = tid ? __frcp_rn((float)tid) : 0.f;
        t.q255d[tid] = __ddiv_rn((double)tid, 255.0);
        const double fh = __ddiv_rn(__dmul_rn((double)tid, 6.0), 255.0), fi = floor(fh);
        t.frac[tid] = __dsub_rn(fh, fi);
        t.sect[tid] = (unsigned char)((int)fi % 6);
    }
}

// a / b correctly rounded for small non-negative integers a <= b (as floats), from r = RN(1/b): q = RN(a r) is within
// one ulp, the remainder a - q b is exact in an fma, and one correction step lands on the correctly rounded quotient
// (Markstein); checked against __fdiv_rn for all 256 x 256 pairs by the all-colour test.
TDL_DEV float div_small(float a, float b, float r) {
    const float q = __fmul_rn(a, r);
    return fmaf(fmaf(-q, b, a), r, q);
}
TDL_DEV double div_const(double a, double b, double r) {      // the same for a double numerator and a constant divisor
    const double q = __dmul_rn(a, r);
    return fma(fma(-q, b, a), r, q);
}

// Pillow Convert.c rgb2hsv_row / hsv2rgb_row around torchvision's wrapping byte addition on H
TDL_DEV Rgb hue_shift(const Rgb& c, int shift, const Tables& T) {
    const int maxc = max(c.r, max(c.g, c.b)), minc = min(c.r, min(c.g, c.b));
    int uh = 0, us = 0;
    const int uv = maxc;
    if (minc != maxc) {
        const float cr = (float)(maxc - minc), rcr = T.rcpf[maxc - minc];
        const float s = div_small(cr, (float)maxc, T.rcpf[maxc]);
        const float rc = div_small((float)(maxc - c.r), cr, rcr), gc = div_small((float)(maxc - c.g), cr, rcr),
                    bc = div_small((float)(maxc - c.b), cr, rcr);
        float h;
        if (c.r == maxc) h = __fsub_rn(bc, gc);
        else if (c.g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
        else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
        // fmod(h / 6.0 + 1.0, 1.0): the argument lies in [5/6, 11/6), so the remainder is an exact subtraction
        double x = __dadd_rn(div_const((double)h, 6.0, 1.0 / 6.0), 1.0);
        if (x >= 1.0) x = __dsub_rn(x, 1.0);
        h = (float)x;
        uh = clip8((int)__dmul_rn((double)h, 255.0));
        us = clip8((int)__dmul_rn((double)s, 255.0));
    }
    uh = (uh + shift) & 255;
    if (us == 0) return Rgb{uv, uv, uv};
    const double fs = T.q255d[us], f = T.frac[uh];
    const double v = (double)uv;
    // (round(): the products never sit on a tie -- oracle/input_pipeline.py's half-even rounding reproduces Pillow's
    //  half-away rounding for all 2^24 HSV triples -- so the one-instruction conversion is exact here)
    const int p = clip8(__double2int_rn(__dmul_rn(v, __dsub_rn(1.0, fs))));
    const int q = clip8(__double2int_rn(__dmul_rn(v, __dsub_rn(1.0, __dmul_rn(fs, f)))));
    const int t = clip8(__double2int_rn(__dmul_rn(v, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f))))));
    switch (T.sect[uh]) {
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
TDL_DEV Rgb apply_chain(Rgb c, const Jitter& j, int stop, int mean, const Tables& T) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= stop) break;
        const int fn = j.order[k];
        if (fn == 0) c = blend3(Rgb{0, 0, 0}, c, j.bright);
        else if (fn == 1) c = blend3(Rgb{mean, mean, mean}, c, j.contrast);
        else if (fn == 2) {
            const int g = gray_u8(c);
            c = blend3(Rgb{g, g, g}, c, j.sat);
        } else if (fn == 3) c = hue_shift(c, j.hue, T);
    }
    return c;
}

// pixel index -> (row, column) without an integer division: y = floor(i * floor(2^32 / W) / 2^32) is the row or one less
TDL_DEV void row_col(int i, int W, unsigned magic, int& y, int& x) {
    y = (int)__umulhi((unsigned)i, magic);
    x = i - y * W;
    if (x >= W) {
        x -= W;
        ++y;
    }
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
    __shared__ Tables T;
    fill_tables(T, threadIdx.x);
    __syncthreads();
    const Jitter j = load_jitter(p, b, fr);
    int cpos = 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (j.order[k] == 1) cpos = min(cpos, k);
    if (cpos == 4) return;                                  // no contrast step in this order (CTA-uniform)
    const int n = p.H * p.W;
    const unsigned magic = p.W > 1 ? (unsigned)(0x100000000ull / (unsigned)p.W) : 0xffffffffu;
    unsigned int acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int y, x;
        row_col(i, p.W, magic, y, x);
        const Rgb c = apply_chain(load_px(p, fr, b, y, x, false), j, cpos, 0, T);     // (the sum does not depend on the flip)
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
    __shared__ Tables T;
    fill_tables(T, threadIdx.x);
    __syncthreads();
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
    // the image's erase boxes, once per CTA (tdl_input_fwd caps erase_count at kMaxBoxes)
    __shared__ int s_box[2 * kMaxBoxes];
    if (mk && (int)threadIdx.x < 2 * p.erase_count) s_box[threadIdx.x] = __ldg(p.holes + (size_t)b * p.erase_count * 2 + threadIdx.x);
    if (mk) __syncthreads();                               // CTA-uniform
    const unsigned magic = p.W > 1 ? (unsigned)(0x100000000ull / (unsigned)p.W) : 0xffffffffu;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int y, x;
        row_col(i, p.W, magic, y, x);
        const Rgb c = load_px(p, fr, b, y, x, flip);
        const float r = T.q255f[c.r], g = T.q255f[c.g], bl = T.q255f[c.b];
        if (co) {
            __stcs(co + i, r);
            __stcs(co + plane + i, g);
            __stcs(co + 2 * plane + i, bl);
        }
        if (ca) {
            if (aug) {
                const Rgb a = apply_chain(c, j, 4, mean, T);
                __stcs(ca + i, T.q255f[a.r]);
                __stcs(ca + plane + i, T.q255f[a.g]);
                __stcs(ca + 2 * plane + i, T.q255f[a.b]);
            } else {
                __stcs(ca + i, r);
                __stcs(ca + plane + i, g);
                __stcs(ca + 2 * plane + i, bl);
            }
        }
        if (mk) {
            float m = 1.f;
            for (int k = 0; k < p.erase_count; ++k)
                if ((unsigned)(y - s_box[2 * k]) < (unsigned)p.erase_h && (unsigned)(x - s_box[2 * k + 1]) < (unsigned)p.erase_w) m = 0.f;
            __stcs(mk + i, m);
            __stcs(mk + plane + i, m);
            __stcs(mk + 2 * plane + i, m);
        }
    }
}

}  // namespace

// grid.x: enough CTAs for ~6 per SM over the (frame, image) planes, each thread looping over several pixels -- the
// per-CTA table fill (256 IEEE divisions) is then amortised over ~13 pixels per thread on the training shape
static int input_grid_x(const InputDev& p) {
    const int n = p.H * p.W, planes = p.nframes * p.B;
    return max(1, min((n + 255) / 256, (148 * 6 + planes - 1) / planes));
}

cudaError_t launch_input_stat(const InputDev& p, cudaStream_t st) {
    dim3 grid(input_grid_x(p), p.nframes, p.B);
    input_stat_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_input_apply(const InputDev& p, cudaStream_t st) {
    dim3 grid(input_grid_x(p), p.nframes, p.B);
    input_apply_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace tdl
