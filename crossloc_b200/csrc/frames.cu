// Input frames: PNG decode (host threads) and the dataset's Resize on the device.
//
// Replaces the per-frame image path of the reference's data loader (/root/reference/dataloader/dataloader.py:306-346:
// skimage.io.imread -> gray2rgb / RGBA -> RGB, and :189-212: transforms.ToPILImage() -> Resize(image_height) -> ToTensor()
// [-> Normalize]) for an evaluation that runs at >10 000 frames/s: the <= 6 DataLoader workers of
// utils/evaluation.py:74 cannot feed that.
//   * cl_decode_png: one PNG file image (8-bit gray / gray+alpha / RGB / RGBA / palette, non-interlaced) -> uint8 HWC RGB,
//     alpha dropped and gray replicated exactly as dataloader.py:312-316 does.  Inflate is zlib's; chunk parsing, the five
//     scanline filters and the colour conversion are done here.  cl_decode_png_batch decodes a batch on host threads.
//   * cl_resize_frames: torchvision's Resize on a PIL image = Pillow's ImagingResample with the bilinear (triangle)
//     filter, antialiased: two separable passes in 22-bit fixed point with a uint8 intermediate (Pillow
//     src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc).
//     The coefficient tables are built on the host in double precision exactly as Pillow builds them, the passes run
//     on the device: bit-identical to Image.resize(..., BILINEAR) (tests/test_frames_*.py).
// ToTensor / Normalize stay in cl_frames_to_nchw (cnn_pointwise.cu).
#include "../../include/crossloc_b200.h"

#include <zlib.h>

#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "cabi_common.h"

namespace cl {
namespace {

// ------------------------------------------------------------------------------------------- PNG
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline uint8_t paeth(int a, int b, int c)
{
    const int p = a + b - c;
    const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (uint8_t)((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c));
}

struct PngInfo {
    int width = 0, height = 0, channels = 0;   // channels of the file: 1 gray, 2 gray+alpha, 3 RGB, 4 RGBA, -1 palette
    std::vector<uint8_t> palette;
    std::vector<uint8_t> idat;
};

// nullptr on success
const char* png_parse(const uint8_t* file, size_t n, PngInfo& info)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n < 8 + 25 || memcmp(file, sig, 8) != 0) return "not a PNG file";
    size_t pos = 8;
    bool have_ihdr = false, done = false;
    while (!done && pos + 12 <= n) {
        const uint32_t len = be32(file + pos);
        const uint8_t* type = file + pos + 4;
        const uint8_t* data = file + pos + 8;
        if ((size_t)len > n - pos - 12) return "truncated PNG chunk";
        if (memcmp(type, "IHDR", 4) == 0) {
            if (len != 13) return "bad IHDR";
            info.width = (int)be32(data);
            info.height = (int)be32(data + 4);
            const int depth = data[8], ctype = data[9], interlace = data[12];
            if (info.width <= 0 || info.height <= 0) return "empty PNG";
            if (depth != 8) return "only 8-bit PNG samples are supported (the datasets of the reference are 8-bit)";
            if (interlace != 0) return "interlaced PNG is not supported";
            if (data[10] != 0 || data[11] != 0) return "unknown PNG compression / filter method";
            switch (ctype) {
                case 0: info.channels = 1; break;
                case 2: info.channels = 3; break;
                case 3: info.channels = -1; break;
                case 4: info.channels = 2; break;
                case 6: info.channels = 4; break;
                default: return "unknown PNG colour type";
            }
            have_ihdr = true;
        } else if (memcmp(type, "PLTE", 4) == 0) {
            info.palette.assign(data, data + len);
        } else if (memcmp(type, "IDAT", 4) == 0) {
            info.idat.insert(info.idat.end(), data, data + len);
        } else if (memcmp(type, "IEND", 4) == 0) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr) return "PNG without IHDR";
    if (info.idat.empty()) return "PNG without image data";
    if (info.channels == -1 && info.palette.size() < 3) return "palette PNG without PLTE";
    return nullptr;
}

// decodes into out [height][width][3]; expects the caller to know the size (png_parse)
const char* png_decode_into(const PngInfo& info, uint8_t* out)
{
    const int bpp = info.channels == -1 ? 1 : info.channels;   // bytes per pixel of the file
    const size_t stride = (size_t)info.width * bpp;
    std::vector<uint8_t> raw((stride + 1) * info.height);
    uLongf raw_len = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &raw_len, info.idat.data(), (uLong)info.idat.size());
    if (zr != Z_OK || raw_len != raw.size()) return "PNG image data does not inflate to the declared size";
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (int y = 0; y < info.height; y++) {
        const uint8_t* line = raw.data() + (stride + 1) * y;
        const int filter = line[0];
        const uint8_t* src = line + 1;
        switch (filter) {
            case 0: memcpy(cur.data(), src, stride); break;
            case 1:
                for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(src[i] + (i >= (size_t)bpp ? cur[i - bpp] : 0));
                break;
            case 2:
                for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(src[i] + prev[i]);
                break;
            case 3:
                for (size_t i = 0; i < stride; i++)
                    cur[i] = (uint8_t)(src[i] + (((i >= (size_t)bpp ? cur[i - bpp] : 0) + prev[i]) >> 1));
                break;
            case 4:
                for (size_t i = 0; i < stride; i++) {
                    const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                    cur[i] = (uint8_t)(src[i] + paeth(a, b, c));
                }
                break;
            default: return "unknown PNG scanline filter";
        }
        uint8_t* dst = out + (size_t)y * info.width * 3;
        switch (info.channels) {
            case 3: memcpy(dst, cur.data(), stride); break;
            case 4:   // RGBA -> RGB (dataloader.py:314-316)
                for (int x = 0; x < info.width; x++) { dst[3 * x] = cur[4 * x]; dst[3 * x + 1] = cur[4 * x + 1]; dst[3 * x + 2] = cur[4 * x + 2]; }
                break;
            case 1:   // gray2rgb (dataloader.py:312-313)
                for (int x = 0; x < info.width; x++) dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = cur[x];
                break;
            case 2:
                for (int x = 0; x < info.width; x++) dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = cur[2 * x];
                break;
            default:  // palette
                for (int x = 0; x < info.width; x++) {
                    const size_t e = (size_t)cur[x] * 3;
                    if (e + 2 >= info.palette.size()) return "PNG palette index out of range";
                    dst[3 * x] = info.palette[e]; dst[3 * x + 1] = info.palette[e + 1]; dst[3 * x + 2] = info.palette[e + 2];
                }
                break;
        }
        prev.swap(cur);
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------- resize (Pillow's resample)
constexpr int kPrecisionBits = 32 - 8 - 2;

struct Coeffs {
    int ksize = 0;
    std::vector<int> bounds;   // [out][2]: first input index, count
    std::vector<int> kk;       // [out][ksize] fixed point
};

// Pillow Resample.c: precompute_coeffs (bilinear filter, support 1) + normalize_coeffs_8bpc
Coeffs pillow_coeffs(int in_size, int out_size)
{
    Coeffs c;
    const double scale = (double)in_size / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;
    c.ksize = (int)ceil(support) * 2 + 1;
    c.bounds.resize((size_t)out_size * 2);
    c.kk.assign((size_t)out_size * c.ksize, 0);
    std::vector<double> k(c.ksize);
    for (int xx = 0; xx < out_size; xx++) {
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; x++) {
            double v = (x + xmin - center + 0.5) * ss;
            if (v < 0.0) v = -v;
            const double w = v < 1.0 ? 1.0 - v : 0.0;
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; x++)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < xmax; x++) {
            const double v = k[x];
            c.kk[(size_t)xx * c.ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
        }
        c.bounds[2 * xx] = xmin;
        c.bounds[2 * xx + 1] = xmax;
    }
    return c;
}

__device__ __forceinline__ uint8_t clip8(int v)
{
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[b][y][xx][c] = clip8(ss + sum_x in[b][y][xmin + x][c] * k[xx][x]);  one thread per output pixel (3 channels)
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int B, int H, int Win,
                                                       int Wout, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize)
{
    const size_t total = (size_t)B * H * Wout;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % Wout);
        const size_t row = i / Wout;   // b * H + y
        const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
        const int* k = kk + (size_t)xx * ksize;
        const uint8_t* src = in + (row * Win + xmin) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < xmax; x++) {
            const int w = k[x];
            s0 += src[3 * x] * w;
            s1 += src[3 * x + 1] * w;
            s2 += src[3 * x + 2] * w;
        }
        uint8_t* dst = out + i * 3;
        dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
    }
}

// out[b][yy][x][c] = clip8(ss + sum_y in[b][ymin + y][x][c] * k[yy][y])
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int B, int Hin, int Hout,
                                                       int W, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize)
{
    const size_t total = (size_t)B * Hout * W * 3;
    const size_t line = (size_t)W * 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t col = i % line;
        const size_t r = i / line;
        const int yy = (int)(r % Hout);
        const size_t b = r / Hout;
        const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
        const int* k = kk + (size_t)yy * ksize;
        const uint8_t* src = in + (b * Hin + ymin) * line + col;
        int s = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < ymax; y++) s += src[(size_t)y * line] * k[y];
        out[i] = clip8(s);
    }
}

}  // namespace
}  // namespace cl

// ------------------------------------------------------------------------------------------- C ABI
extern "C" int cl_png_info(const void* file, size_t file_bytes, int* height, int* width, int* file_channels)
{
    using namespace cl;
    if (!file || !height || !width) return fail(-1, "cl_png_info: NULL argument");
    PngInfo info;
    if (const char* e = png_parse(static_cast<const uint8_t*>(file), file_bytes, info)) return fail(-1, "cl_png_info: %s", e);
    *height = info.height;
    *width = info.width;
    if (file_channels) *file_channels = info.channels == -1 ? 3 : info.channels;
    return 0;
}

extern "C" int cl_decode_png(const void* file, size_t file_bytes, uint8_t* out_rgb, int height, int width)
{
    using namespace cl;
    if (!file || !out_rgb) return fail(-1, "cl_decode_png: NULL argument");
    PngInfo info;
    if (const char* e = png_parse(static_cast<const uint8_t*>(file), file_bytes, info)) return fail(-1, "cl_decode_png: %s", e);
    if (info.height != height || info.width != width)
        return fail(-1, "cl_decode_png: the file holds a %dx%d image, the destination is %dx%d", info.height, info.width, height, width);
    if (const char* e = png_decode_into(info, out_rgb)) return fail(-1, "cl_decode_png: %s", e);
    return 0;
}

extern "C" int cl_decode_png_batch(const void* const* files, const size_t* file_bytes, int count, uint8_t* out_rgb, int height,
                                   int width, int threads)
{
    using namespace cl;
    if (!files || !file_bytes || !out_rgb || count < 0) return fail(-1, "cl_decode_png_batch: invalid argument");
    if (threads < 1) threads = 1;
    if (threads > count) threads = count > 0 ? count : 1;
    std::vector<std::string> errors((size_t)threads);
    auto work = [&](int t) {
        for (int i = t; i < count; i += threads) {
            PngInfo info;
            const char* e = png_parse(static_cast<const uint8_t*>(files[i]), file_bytes[i], info);
            if (!e && (info.height != height || info.width != width)) e = "frame size differs from the batch size";
            if (!e) e = png_decode_into(info, out_rgb + (size_t)i * height * width * 3);
            if (e && errors[t].empty()) errors[t] = "frame " + std::to_string(i) + ": " + e;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (const auto& e : errors)
        if (!e.empty()) return fail(-1, "cl_decode_png_batch: %s", e.c_str());
    return 0;
}

extern "C" int cl_resize_frames(const uint8_t* src, int B, int Hin, int Win, uint8_t* dst, int Hout, int Wout, void* workspace,
                                size_t workspace_bytes, void* cuda_stream)
{
    using namespace cl;
    static const char* kFn = "cl_resize_frames";
    if (!src || !dst || !is_device_ptr(src) || !is_device_ptr(dst)) return fail(-1, "%s: src and dst must be device pointers", kFn);
    if (B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return fail(-1, "%s: invalid sizes", kFn);
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    if (Hin == Hout && Win == Wout) {   // Image.resize returns a copy when the size does not change
        CL_CUDA(cudaMemcpyAsync(dst, src, (size_t)B * Hin * Win * 3, cudaMemcpyDeviceToDevice, stream));
        return 0;
    }
    const Coeffs ch = pillow_coeffs(Win, Wout), cv = pillow_coeffs(Hin, Hout);
    // workspace: horizontal result [B][Hin][Wout][3], then the four coefficient tables
    const size_t mid_bytes = ((size_t)B * Hin * Wout * 3 + 255) & ~(size_t)255;
    const size_t tab_ints = ch.bounds.size() + ch.kk.size() + cv.bounds.size() + cv.kk.size();
    const size_t need = mid_bytes + tab_ints * sizeof(int);
    if (!workspace || !is_device_ptr(workspace) || workspace_bytes < need)
        return fail(-3, "%s: workspace of %zu bytes needed (device memory)", kFn, need);
    uint8_t* mid = static_cast<uint8_t*>(workspace);
    int* tab = reinterpret_cast<int*>(mid + mid_bytes);
    std::vector<int> host;
    host.reserve(tab_ints);
    host.insert(host.end(), ch.bounds.begin(), ch.bounds.end());
    host.insert(host.end(), ch.kk.begin(), ch.kk.end());
    host.insert(host.end(), cv.bounds.begin(), cv.bounds.end());
    host.insert(host.end(), cv.kk.begin(), cv.kk.end());
    CL_CUDA(cudaMemcpyAsync(tab, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    CL_CUDA(cudaStreamSynchronize(stream));   // `host` is pageable and leaves scope
    const int* hb = tab;
    const int* hk = hb + ch.bounds.size();
    const int* vb = hk + ch.kk.size();
    const int* vk = vb + cv.bounds.size();
    const uint8_t* vsrc = src;
    if (Win != Wout) {
        const size_t total = (size_t)B * Hin * Wout;
        unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 16 ? (total + 255) / 256 : 148u * 16);
        resize_h_kernel<<<blocks, 256, 0, stream>>>(src, Hin == Hout ? dst : mid, B, Hin, Win, Wout, hb, hk, ch.ksize);
        vsrc = mid;
    }
    if (Hin != Hout) {
        const size_t total = (size_t)B * Hout * Wout * 3;
        unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 32 ? (total + 255) / 256 : 148u * 32);
        resize_v_kernel<<<blocks, 256, 0, stream>>>(vsrc, dst, B, Hin, Hout, Wout, vb, vk, cv.ksize);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-2, "%s: %s", kFn, cudaGetErrorString(e));
    return 0;
}

// The fixed-point coefficient table of one axis (host function): kk [out_size][ksize] and bounds [out_size][2] =
// (first input index, taps).  Pass NULL arrays to query ksize.  What cl_resize_frames uploads; exposed so that the
// table can be checked against Pillow without a device.
extern "C" int cl_resize_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk)
{
    using namespace cl;
    if (in_size <= 0 || out_size <= 0 || !ksize) return fail(-1, "cl_resize_coeffs: invalid argument");
    const Coeffs c = pillow_coeffs(in_size, out_size);
    *ksize = c.ksize;
    if (bounds) memcpy(bounds, c.bounds.data(), c.bounds.size() * sizeof(int));
    if (kk) memcpy(kk, c.kk.data(), c.kk.size() * sizeof(int));
    return 0;
}

// bytes of device workspace cl_resize_frames needs for these sizes
extern "C" size_t cl_resize_workspace_bytes(int B, int Hin, int Win, int Hout, int Wout)
{
    using namespace cl;
    if (B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return 0;
    auto ks = [](int in, int out) {
        double fs = (double)in / out;
        if (fs < 1.0) fs = 1.0;
        return (size_t)((int)ceil(fs) * 2 + 1);
    };
    const size_t mid = ((size_t)B * Hin * Wout * 3 + 255) & ~(size_t)255;
    return mid + ((size_t)Wout * (2 + ks(Win, Wout)) + (size_t)Hout * (2 + ks(Hin, Hout))) * sizeof(int);
}
