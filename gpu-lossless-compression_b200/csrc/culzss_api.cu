// Reference-named CULZSS entry points (include/culzss_gpu.h) on top of the batch kernels.
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "../../include/b200lc.h"
#include "../../include/culzss_gpu.h"

using namespace b200lc;

namespace {
constexpr int kGroups = 4;          // the reference keeps 4 groups of 16 streams (gpu_compress.cu:82-84)
cudaStream_t g_streams[kGroups] = {nullptr, nullptr, nullptr, nullptr};
void *g_scratch[kGroups] = {nullptr, nullptr, nullptr, nullptr};
size_t g_scratch_bytes[kGroups] = {0, 0, 0, 0};
u32 *g_len_d[kGroups] = {nullptr, nullptr, nullptr, nullptr};
bool g_init = false;
std::mutex g_mu;

// decode side (one call at a time, like the reference's single degpu_consumer thread)
void *g_dec_in = nullptr, *g_dec_out = nullptr, *g_dec_scratch = nullptr;
size_t g_dec_in_bytes = 0, g_dec_out_bytes = 0, g_dec_scratch_bytes = 0;
u64 *g_dec_off = nullptr;
cudaStream_t g_dec_stream = nullptr;

bool ok(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "b200lc culzss: %s: %s\n", what, cudaGetErrorString(e));
        return false;
    }
    return true;
}
void ensure_init()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_init) return;
    for (int i = 0; i < kGroups; ++i) {
        ok(cudaStreamCreateWithFlags(&g_streams[i], cudaStreamNonBlocking), "stream create");
        ok(cudaMalloc(&g_len_d[i], 256), "cudaMalloc");
    }
    g_init = true;
}
bool grow(void **p, size_t *have, size_t need)
{
    if (*have >= need) return true;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    if (!ok(cudaMalloc(p, need), "cudaMalloc")) return false;
    *have = need;
    return true;
}
}  // namespace

extern "C" unsigned char *initGPUmem(int buf_length)
{
    unsigned char *p = nullptr;
    ok(cudaMalloc(&p, (size_t)buf_length), "initGPUmem");
    return p;
}
extern "C" unsigned char *initCPUmem(int buf_length)
{
    unsigned char *p = nullptr;
    ok(cudaMallocHost(&p, (size_t)buf_length), "initCPUmem");
    return p;
}
extern "C" void deleteGPUmem(unsigned char *mem_d) { cudaFree(mem_d); }
extern "C" void deleteCPUmem(unsigned char *mem_d) { cudaFreeHost(mem_d); }
extern "C" void initGPU(void) { ensure_init(); }
extern "C" void resetGPU(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    cudaDeviceReset();
    ++context_epoch();          // kernel attributes set in the old context are gone
    g_init = false;
    for (int i = 0; i < kGroups; ++i) {
        g_streams[i] = nullptr;
        g_scratch[i] = nullptr;
        g_scratch_bytes[i] = 0;
        g_len_d[i] = nullptr;
    }
    g_dec_in = g_dec_out = g_dec_scratch = nullptr;
    g_dec_in_bytes = g_dec_out_bytes = g_dec_scratch_bytes = 0;
    g_dec_off = nullptr;
    g_dec_stream = nullptr;
}
extern "C" int streams_in_GPU(void) { return 1; }
extern "C" int onestream_finish_GPU(int index)
{
    ensure_init();
    if (index < 0 || index >= kGroups) index = ((index % kGroups) + kGroups) % kGroups;
    return ok(cudaStreamSynchronize(g_streams[index]), "onestream_finish_GPU") ? 1 : 0;
}
extern "C" void deleteGPUStreams(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_init) return;
    for (int i = 0; i < kGroups; ++i) {
        cudaStreamDestroy(g_streams[i]);
        cudaFree(g_scratch[i]);
        cudaFree(g_len_d[i]);
        g_streams[i] = nullptr;
        g_scratch[i] = nullptr;
        g_scratch_bytes[i] = 0;
        g_len_d[i] = nullptr;
    }
    g_init = false;
}

// B200LC_CULZSS_FAST=1|2|4|lane makes the reference-named wrappers and the container writer use the
// NON-PARITY fast encoder mode (same format, other matches); unset = the reference's bytes.
static int fast_depth()
{
    static int d = 0;
    static bool read = false;
    if (!read) {
        const char *e = getenv("B200LC_CULZSS_FAST");
        const int v = e ? atoi(e) : 0;
        d = (v == 1 || v == 2 || v == 4) ? v : 0;
        if (e && (!strcmp(e, "lane") || v == B200LC_CULZSS_FAST_LANE)) d = B200LC_CULZSS_FAST_LANE;
        read = true;
    }
    return d;
}
static int encode_any(const uint8_t *d_in, size_t nbuf, size_t buf_length, uint8_t *d_out, size_t stride,
                      uint32_t *d_len, void *scratch, size_t scratch_bytes, cudaStream_t st)
{
    const int d = fast_depth();
    return d ? b200lc_culzss_encode_fast_batch(d_in, nbuf, buf_length, d_out, stride, d_len, scratch, scratch_bytes, d, st)
             : b200lc_culzss_encode_batch(d_in, nbuf, buf_length, d_out, stride, d_len, scratch, scratch_bytes, st);
}

// bufferout layout handed from compression_kernel_wrapper to aftercompression_wrapper:
//   [0,4) compressed size incl. trailer, 0 = "compression took more"; [16, 16 + size) the buffer.
extern "C" int compression_kernel_wrapper(unsigned char *buffer, int buf_length,
                                          unsigned char *bufferout, int, int, int, int, int index,
                                          unsigned char *in_d, unsigned char *out_d)
{
    ensure_init();
    if (!buffer || !bufferout || !in_d || !out_d || buf_length <= 0 || buf_length % 4096) {
        fprintf(stderr, "b200lc culzss: compression_kernel_wrapper: bad arguments\n");
        return 0;
    }
    if (index < 0 || index >= kGroups) index = ((index % kGroups) + kGroups) % kGroups;
    cudaStream_t st = g_streams[index];
    const size_t need = b200lc_culzss_encode_scratch_bytes(1, (size_t)buf_length);
    if (!grow(&g_scratch[index], &g_scratch_bytes[index], need)) return 0;
    if (!ok(cudaMemcpyAsync(in_d, buffer, (size_t)buf_length, cudaMemcpyHostToDevice, st), "H2D")) return 0;
    // out_d holds 2 * buf_length bytes: more than the worst-case compressed buffer
    const size_t stride = 2 * (size_t)buf_length;
    if (encode_any(in_d, 1, (size_t)buf_length, out_d, stride, g_len_d[index],
                                   g_scratch[index], g_scratch_bytes[index], st) != B200LC_OK)
        return 0;
    const size_t max_comp = (size_t)buf_length + 2 * ((size_t)buf_length / 4096) + 6 + 32;
    const size_t copy = max_comp + 16 <= stride ? max_comp : stride - 16;
    if (!ok(cudaMemcpyAsync(bufferout, g_len_d[index], 4, cudaMemcpyDeviceToHost, st), "D2H")) return 0;
    if (!ok(cudaMemcpyAsync(bufferout + 16, out_d, copy, cudaMemcpyDeviceToHost, st), "D2H")) return 0;
    return 1;
}

extern "C" int aftercompression_wrapper(unsigned char *buffer, int buf_length,
                                        unsigned char *bufferout, int *comp_length)
{
    (void)buf_length;
    u32 len = 0;
    memcpy(&len, bufferout, 4);
    if (len == 0) {
        printf("compression took more!!! \n");
        return 0;
    }
    memcpy(buffer, bufferout + 16, len);
    *comp_length = (int)len;
    return 1;
}

extern "C" unsigned char *deinitGPUmem(int buf_length) { return initGPUmem(buf_length); }
extern "C" void dedeleteGPUmem(unsigned char *mem_d) { cudaFree(mem_d); }
extern "C" void deinitGPU(void) { ensure_init(); }

extern "C" int decompression_kernel_wrapper(unsigned char *buffer, int buf_length,
                                            int *decomp_length, int, int, int)
{
    ensure_init();
    if (!buffer || buf_length < 6 || !decomp_length) return 0;
    // trailer (gpu_decompress.cu:258-270)
    const u32 orig = ((u32)buffer[buf_length - 6] << 24) | ((u32)buffer[buf_length - 5] << 16) |
                     ((u32)buffer[buf_length - 4] << 8) | (u32)buffer[buf_length - 3];
    const u32 pad = ((u32)buffer[buf_length - 2] << 8) | (u32)buffer[buf_length - 1];
    if (orig == 0 || orig % 4096 || pad > orig) {
        fprintf(stderr, "b200lc culzss: malformed trailer\n");
        return 0;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_dec_stream) ok(cudaStreamCreateWithFlags(&g_dec_stream, cudaStreamNonBlocking), "stream");
    if (!g_dec_off) ok(cudaMalloc(&g_dec_off, 64), "cudaMalloc");
    if (!grow(&g_dec_in, &g_dec_in_bytes, (size_t)buf_length + 64)) return 0;
    if (!grow(&g_dec_out, &g_dec_out_bytes, (size_t)orig)) return 0;
    const size_t need = b200lc_culzss_decode_scratch_bytes(1, orig);
    if (!grow(&g_dec_scratch, &g_dec_scratch_bytes, need)) return 0;
    const u64 offs[2] = {0, (u64)buf_length};
    cudaStream_t st = g_dec_stream;
    if (!ok(cudaMemcpyAsync(g_dec_off, offs, sizeof(offs), cudaMemcpyHostToDevice, st), "H2D")) return 0;
    if (!ok(cudaMemcpyAsync(g_dec_in, buffer, (size_t)buf_length, cudaMemcpyHostToDevice, st), "H2D")) return 0;
    // a compressed size equal to the original size would be taken for a raw buffer by the batch
    // API; the reference never calls this wrapper for raw buffers (deculzss.c:94-95)
    if (b200lc_culzss_decode_batch((const u8 *)g_dec_in, g_dec_off, 1, orig, (u8 *)g_dec_out,
                                   g_dec_scratch, g_dec_scratch_bytes, st) != B200LC_OK)
        return 0;
    if (!ok(cudaMemcpyAsync(buffer, g_dec_out, orig - pad, cudaMemcpyDeviceToHost, st), "D2H")) return 0;
    if (!ok(cudaStreamSynchronize(st), "sync")) return 0;
    *decomp_length = (int)(orig - pad);
    return 1;
}

// ============================================================================ file container
// Format written by the reference CLI (cuda-lzss-cluster/main.c:236-245, culzss.c:220,243-264,
// decompression.c:90-141): native-endian u32 nblocks, u32 padding, u32 cumulative_end[nblocks],
// then the buffers; a buffer whose stored size equals the 1 MiB buffer size is raw.  The reference
// pads a last partial buffer with stale bytes of the previous one (main.c:122-130);
// here the padding is zero, so containers of inputs that are not a multiple of 1 MiB decode
// correctly everywhere but are not byte-identical to the reference's.
namespace {
constexpr size_t kBuf = 1u << 20;   // BUFSIZE, main.c:62
}

extern "C" size_t b200lc_culzss_container_bound(size_t n)
{
    const size_t nb = (n + kBuf - 1) / kBuf;
    return 8 + 4 * nb + nb * (kBuf + 2 * (kBuf / 4096) + 6 + 32);
}

// ------------------------------------------------------------------------------ file container
// Work area of the container functions, kept between calls (cudaMalloc / cudaFree of gigabytes per
// call cost more than the kernels) and guarded by a mutex; it follows the caller's current device.
namespace {
struct ContainerArea {
    int device = -1;
    unsigned epoch = 0;
    void *in = nullptr, *out = nullptr, *scratch = nullptr, *compact = nullptr, *small = nullptr;
    size_t in_b = 0, out_b = 0, scratch_b = 0, compact_b = 0, small_b = 0;
    void drop()
    {
        cudaFree(in); cudaFree(out); cudaFree(scratch); cudaFree(compact); cudaFree(small);
        in = out = scratch = compact = small = nullptr;
        in_b = out_b = scratch_b = compact_b = small_b = 0;
    }
};
ContainerArea g_ct;
std::mutex g_ct_mu;

bool container_area(size_t in_b, size_t out_b, size_t scratch_b, size_t compact_b, size_t small_b)
{
    int dev = 0;
    if (!ok(cudaGetDevice(&dev), "cudaGetDevice")) return false;
    if (g_ct.epoch != context_epoch()) {        // the context was reset: the old pointers are gone
        g_ct = ContainerArea();
        g_ct.epoch = context_epoch();
    }
    if (g_ct.device != dev) {
        if (g_ct.device >= 0) {
            cudaSetDevice(g_ct.device);
            g_ct.drop();
            cudaSetDevice(dev);
        }
        g_ct.device = dev;
    }
    return grow(&g_ct.in, &g_ct.in_b, in_b) && grow(&g_ct.out, &g_ct.out_b, out_b) &&
           grow(&g_ct.scratch, &g_ct.scratch_b, scratch_b) && grow(&g_ct.compact, &g_ct.compact_b, compact_b) &&
           grow(&g_ct.small, &g_ct.small_b, small_b);
}

// buffer b of the container = its compressed bytes, or the raw input buffer when len[b] == 0
// ("compression took more", culzss.c:177-183,241-242), at byte offset offs[b] of dst
__global__ void __launch_bounds__(256) culzss_compact_kernel(const u8 *__restrict__ comp, size_t stride,
                                                             const u8 *__restrict__ raw, size_t buf,
                                                             const u32 *__restrict__ len,
                                                             const u64 *__restrict__ offs, u8 *__restrict__ dst)
{
    const size_t b = blockIdx.y;
    const u32 l = len[b];
    const size_t sz = l ? l : buf;
    const u8 *src = l ? comp + b * stride : raw + b * buf;
    u8 *d = dst + offs[b];
    // destination words are 4-byte aligned from `head` on; the sources are 16-byte aligned
    const size_t head = min(sz, (size_t)((4 - (reinterpret_cast<uintptr_t>(d) & 3)) & 3));
    const size_t words = (sz - head) >> 2;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, tn = (size_t)gridDim.x * blockDim.x;
    if (t0 < head) d[t0] = src[t0];
    const u32 sh = 8 * (u32)(head & 3);
    const u32 *sw = reinterpret_cast<const u32 *>(src);       // source word i holds bytes 4 i ..
    u32 *dw = reinterpret_cast<u32 *>(d + head);
    for (size_t i = t0; i < words; i += tn) {
        const size_t at = head + 4 * i;                      // source byte offset of this word
        const u32 lo = sw[at >> 2], hi = sh ? sw[(at >> 2) + 1] : 0u;
        dw[i] = __funnelshift_r(lo, hi, sh);
    }
    const size_t done = head + 4 * words;
    if (t0 < sz - done) d[done + t0] = src[done + t0];
}

// Three streams per call: copies up, kernels, copies down.  Created once per device context.
struct ContainerStreams {
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    std::vector<cudaEvent_t> ev;
    unsigned epoch = 0;
    int device = -1;
};
ContainerStreams g_cs;

bool container_streams(size_t events)
{
    int dev = 0;
    if (!ok(cudaGetDevice(&dev), "cudaGetDevice")) return false;
    if (g_cs.epoch != context_epoch() || g_cs.device != dev) {      // new context / device: start over
        g_cs = ContainerStreams();
        g_cs.epoch = context_epoch();
        g_cs.device = dev;
    }
    if (!g_cs.up && (!ok(cudaStreamCreateWithFlags(&g_cs.up, cudaStreamNonBlocking), "stream") ||
                     !ok(cudaStreamCreateWithFlags(&g_cs.run, cudaStreamNonBlocking), "stream") ||
                     !ok(cudaStreamCreateWithFlags(&g_cs.down, cudaStreamNonBlocking), "stream")))
        return false;
    while (g_cs.ev.size() < events) {
        cudaEvent_t e;
        if (!ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")) return false;
        g_cs.ev.push_back(e);
    }
    return true;
}

// buffers per pipeline chunk: large enough for the packet-per-lane encoder (>= 160 MiB per call)
constexpr size_t kChunkBufs = 256;
}  // namespace

// Pipelined over chunks of 256 buffers: while chunk c is being coded, chunk c + 1 travels up and the
// gathered image of chunk c - 1 travels down (pinned host buffers overlap fully; pageable ones are
// staged by the driver).  The header needs the sizes of all buffers, the payload of a chunk only
// those of the chunks before it.
extern "C" int b200lc_culzss_compress_container(const uint8_t *h_in, size_t n, uint8_t *h_out,
                                                size_t cap, size_t *out_len)
{
    if (!h_in || !h_out || !out_len) return B200LC_ERR_ARG;
    if (n < kBuf || n >= (size_t(1) << 32)) return B200LC_ERR_UNSUPPORTED;   // main.c:225-229: "too small"
    const size_t nb = (n + kBuf - 1) / kBuf;
    const size_t padding = nb * kBuf - n;
    const size_t stride = (kBuf + kBuf / 8 + 1024 + 15) & ~size_t(15);
    const size_t cb = std::min(nb, kChunkBufs);                      // buffers per chunk
    const size_t nchunks = (nb + cb - 1) / cb;
    const size_t sb = b200lc_culzss_encode_scratch_bytes(cb, kBuf);
    std::lock_guard<std::mutex> lk(g_ct_mu);
    // small: len[nb] (u32) | offs[nb] (u64)
    const size_t small_b = ((nb * 4 + 255) & ~size_t(255)) + nb * 8;
    if (!container_area(nb * kBuf + 16, cb * stride, sb, nb * kBuf + 64, small_b)) return B200LC_ERR_CUDA;
    if (!container_streams(2 * nchunks)) return B200LC_ERR_CUDA;
    u8 *d_in = static_cast<u8 *>(g_ct.in), *d_out = static_cast<u8 *>(g_ct.out);
    u8 *d_compact = static_cast<u8 *>(g_ct.compact);
    u32 *d_len = static_cast<u32 *>(g_ct.small);
    u64 *d_offs = reinterpret_cast<u64 *>(static_cast<u8 *>(g_ct.small) + ((nb * 4 + 255) & ~size_t(255)));
    const size_t hdr_bytes = 8 + 4 * nb;
    if (hdr_bytes > cap) return B200LC_ERR_OVERFLOW;
    u32 *hdr = reinterpret_cast<u32 *>(h_out);
    hdr[0] = (u32)nb;
    hdr[1] = (u32)padding;

    // everything up, chunk by chunk (the copy engine works through them while the kernels start)
    for (size_t c = 0; c < nchunks; ++c) {
        const size_t lo = c * cb * kBuf, hi = std::min(n, (c + 1) * cb * kBuf);
        if (!ok(cudaMemcpyAsync(d_in + lo, h_in + lo, hi - lo, cudaMemcpyHostToDevice, g_cs.up), "H2D")) return B200LC_ERR_CUDA;
        if (c + 1 == nchunks && !ok(cudaMemsetAsync(d_in + n, 0, padding + 16, g_cs.up), "memset")) return B200LC_ERR_CUDA;
        if (!ok(cudaEventRecord(g_cs.ev[2 * c], g_cs.up), "event")) return B200LC_ERR_CUDA;
    }
    std::vector<u32> len(nb);
    std::vector<u64> offs(nb);
    size_t cum = 0;
    int rc = B200LC_OK;
    for (size_t c = 0; c < nchunks && rc == B200LC_OK; ++c) {
        const size_t b0 = c * cb, nbc = std::min(cb, nb - b0);
        if (!ok(cudaStreamWaitEvent(g_cs.run, g_cs.ev[2 * c], 0), "wait")) { rc = B200LC_ERR_CUDA; break; }
        // the gather of the chunk before must have read d_out before this chunk's encoder rewrites it:
        // both run on g_cs.run, in order
        rc = encode_any(d_in + b0 * kBuf, nbc, kBuf, d_out, stride, d_len + b0, g_ct.scratch, sb, g_cs.run);
        if (rc != B200LC_OK) break;
        if (!ok(cudaMemcpyAsync(len.data() + b0, d_len + b0, nbc * 4, cudaMemcpyDeviceToHost, g_cs.run), "D2H") ||
            !ok(cudaStreamSynchronize(g_cs.run), "sync")) { rc = B200LC_ERR_CUDA; break; }
        const size_t chunk_base = cum;
        for (size_t b = b0; b < b0 + nbc; ++b) {
            offs[b] = cum;
            cum += len[b] ? len[b] : kBuf;
            hdr[2 + b] = (u32)cum;                       // cumulative ends (culzss.c:220,243-264)
        }
        if (hdr_bytes + cum > cap) { rc = B200LC_ERR_OVERFLOW; break; }
        if (!ok(cudaMemcpyAsync(d_offs + b0, offs.data() + b0, nbc * 8, cudaMemcpyHostToDevice, g_cs.run), "H2D")) { rc = B200LC_ERR_CUDA; break; }
        culzss_compact_kernel<<<dim3(16, (unsigned)nbc), 256, 0, g_cs.run>>>(d_out, stride, d_in + b0 * kBuf, kBuf, d_len + b0,
                                                                            d_offs + b0, d_compact);
        if (!ok(cudaGetLastError(), "compact") || !ok(cudaEventRecord(g_cs.ev[2 * c + 1], g_cs.run), "event") ||
            !ok(cudaStreamWaitEvent(g_cs.down, g_cs.ev[2 * c + 1], 0), "wait") ||
            !ok(cudaMemcpyAsync(h_out + hdr_bytes + chunk_base, d_compact + chunk_base, cum - chunk_base,
                                cudaMemcpyDeviceToHost, g_cs.down), "D2H"))
            rc = B200LC_ERR_CUDA;
    }
    // nothing may still be in flight when the buffers go back to the caller (or to the next call)
    const bool drained = ok(cudaStreamSynchronize(g_cs.up), "sync") & ok(cudaStreamSynchronize(g_cs.run), "sync") &
                         ok(cudaStreamSynchronize(g_cs.down), "sync");
    if (rc != B200LC_OK) return rc;
    if (!drained) return B200LC_ERR_CUDA;
    *out_len = hdr_bytes + cum;
    return B200LC_OK;
}

extern "C" int b200lc_culzss_decompress_container(const uint8_t *h_in, size_t n, uint8_t *h_out,
                                                  size_t cap, size_t *out_len)
{
    if (!h_in || !h_out || !out_len || n < 8) return B200LC_ERR_ARG;
    const u32 *hdr = reinterpret_cast<const u32 *>(h_in);
    const size_t nb = hdr[0], padding = hdr[1];
    if (nb == 0 || n < 8 + 4 * nb || padding >= kBuf) return B200LC_ERR_ARG;
    const size_t payload = hdr[2 + nb - 1];
    if (8 + 4 * nb + payload > n) return B200LC_ERR_ARG;
    const size_t out_bytes = nb * kBuf - padding;
    if (out_bytes > cap) return B200LC_ERR_OVERFLOW;
    std::vector<u64> offs(nb + 1);
    offs[0] = 0;
    for (size_t b = 0; b < nb; ++b) {
        offs[b + 1] = hdr[2 + b];
        // cumulative ends must grow, by at most the worst-case size of one stored buffer
        // (aftercomp's test lets the payload pass buf_length by one token group, then the trailer)
        if (offs[b + 1] <= offs[b] || offs[b + 1] - offs[b] > kBuf + 2 * (kBuf / 4096) + 6 + 32)
            return B200LC_ERR_ARG;
    }
    // pipelined like the compressor: chunk c + 1 travels up while chunk c is decoded and chunk c - 1
    // travels down
    const size_t cb = std::min(nb, kChunkBufs);
    const size_t nchunks = (nb + cb - 1) / cb;
    const size_t sb = b200lc_culzss_decode_scratch_bytes(cb, kBuf);
    std::lock_guard<std::mutex> lk(g_ct_mu);
    if (!container_area(payload + 64, nb * kBuf, sb, 0, (nb + nchunks) * 8)) return B200LC_ERR_CUDA;
    if (!container_streams(2 * nchunks)) return B200LC_ERR_CUDA;
    u8 *d_comp = static_cast<u8 *>(g_ct.in), *d_out = static_cast<u8 *>(g_ct.out);
    u64 *d_offs = static_cast<u64 *>(g_ct.small);
    // per chunk its own offset table (nbc + 1 entries, absolute offsets into d_comp)
    std::vector<u64> table(nb + nchunks);
    for (size_t c = 0, t = 0; c < nchunks; ++c) {
        const size_t b0 = c * cb, nbc = std::min(cb, nb - b0);
        for (size_t k = 0; k <= nbc; ++k) table[t++] = offs[b0 + k];
    }
    int rc = B200LC_OK;
    if (!ok(cudaMemcpyAsync(d_offs, table.data(), table.size() * 8, cudaMemcpyHostToDevice, g_cs.up), "H2D")) rc = B200LC_ERR_CUDA;
    for (size_t c = 0; c < nchunks && rc == B200LC_OK; ++c) {
        const size_t b0 = c * cb, nbc = std::min(cb, nb - b0);
        const size_t lo = offs[b0], hi = offs[b0 + nbc];
        if (!ok(cudaMemcpyAsync(d_comp + lo, h_in + 8 + 4 * nb + lo, hi - lo, cudaMemcpyHostToDevice, g_cs.up), "H2D") ||
            !ok(cudaEventRecord(g_cs.ev[2 * c], g_cs.up), "event"))
            rc = B200LC_ERR_CUDA;
    }
    for (size_t c = 0; c < nchunks && rc == B200LC_OK; ++c) {
        const size_t b0 = c * cb, nbc = std::min(cb, nb - b0);
        if (!ok(cudaStreamWaitEvent(g_cs.run, g_cs.ev[2 * c], 0), "wait")) { rc = B200LC_ERR_CUDA; break; }
        rc = b200lc_culzss_decode_batch(d_comp, d_offs + b0 + c, nbc, kBuf, d_out + b0 * kBuf, g_ct.scratch, sb, g_cs.run);
        if (rc != B200LC_OK) break;
        const size_t out_lo = b0 * kBuf, out_hi = std::min(out_bytes, (b0 + nbc) * kBuf);
        if (!ok(cudaEventRecord(g_cs.ev[2 * c + 1], g_cs.run), "event") ||
            !ok(cudaStreamWaitEvent(g_cs.down, g_cs.ev[2 * c + 1], 0), "wait") ||
            !ok(cudaMemcpyAsync(h_out + out_lo, d_out + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, g_cs.down), "D2H"))
            rc = B200LC_ERR_CUDA;
    }
    const bool drained = ok(cudaStreamSynchronize(g_cs.up), "sync") & ok(cudaStreamSynchronize(g_cs.run), "sync") &
                         ok(cudaStreamSynchronize(g_cs.down), "sync");
    if (rc != B200LC_OK) return rc;
    if (!drained) return B200LC_ERR_CUDA;
    *out_len = out_bytes;
    return B200LC_OK;
}
