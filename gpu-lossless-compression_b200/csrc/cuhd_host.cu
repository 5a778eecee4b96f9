// Host-buffer entry points for the CUHD path: what the reference's demo does around
// CUHDGPUDecoder::decode (cuhd-icpp/src/demo.cc:122-168: allocate device buffers, H2D of table
// and stream, decode, D2H of the symbols) and around the CPU encoder (demo.cc:90-107), as one
// C-ABI session object that owns the device buffers, the scratch and a stream.
//
// Host pointers handed to these functions should be pinned (cudaHostAlloc / cudaHostRegister);
// pageable memory works but copies are then staged by the driver.
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "../../include/b200lc.h"

using namespace b200lc;

struct b200lc_cuhd_session {
    int device = 0;          // the device the session was created on; every call runs there
    size_t max_symbols = 0;
    size_t max_units = 0;
    cudaStream_t stream = nullptr;
    u8 *d_symbols = nullptr;
    u32 *d_units = nullptr;
    u8 *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    u32 *d_piece_hist = nullptr;   // one 256-bin histogram per encoder piece
    u8 *d_small = nullptr;   // [0,2048) hist | [2048,3072) code | [3072,3328) len | [4096,..) lut | total_bits
    u64 *h_small = nullptr;  // pinned: hist[256] + total_bits + kMaxChunks progress words
    cudaStream_t h2d = nullptr, d2h = nullptr;   // copy streams of the pipelined decode
    std::vector<cudaEvent_t> events;
};

// A session may be driven from any host thread (a fresh thread's current device is 0, not the
// session's): run the call on the session's device and put the caller's device back afterwards.
struct DeviceScope {
    int prev = -1;
    bool switched = false;
    explicit DeviceScope(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceScope()
    {
        if (switched) cudaSetDevice(prev);
    }
};

static const size_t kMaxChunks = 512;
// Chunks of the pipelined copies: a sixteenth of the transfer, between 8 and 32 MiB of stream for
// the decoder and between 16 and 64 MiB of symbols for the encoder.  Measured on C2 with two encode
// and two decode sessions in flight (bench.py e2e): 4 / 16 MiB chunks 23.2 GB/s, 8 / 32 MiB 24.9,
// 16 / 32 MiB 26.2, 32 / 64 MiB 26.9, 64 / 128 MiB 27.2, 128 / 128 MiB 27.0 -- many small copies
// from four streams leave gaps on the two copy engines.  B200LC_CUHD_CHUNK_MB /
// B200LC_CUHD_ENC_CHUNK_MB fix the sizes (tuning).
static size_t env_mb(const char *name)
{
    const char *e = getenv(name);
    const long v = e ? atol(e) : 0;
    return v >= 1 && v <= 1024 ? (size_t)v << 20 : 0;
}
static size_t chunk_bytes(size_t total, size_t lo, size_t hi, bool encoder)
{
    static const size_t fixed_dec = env_mb("B200LC_CUHD_CHUNK_MB"), fixed_enc = env_mb("B200LC_CUHD_ENC_CHUNK_MB");
    const size_t fixed = encoder ? fixed_enc : fixed_dec;
    if (fixed) return fixed;
    const size_t c = (total / 16 + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    return c < lo ? lo : (c > hi ? hi : c);
}

static const size_t kSmallBytes = 4096 + (size_t(2) << 13) + 64;

extern "C" int b200lc_cuhd_session_create(size_t max_symbols, b200lc_cuhd_session **out)
{
    if (!out || max_symbols == 0) return B200LC_ERR_ARG;
    b200lc_cuhd_session *s = new (std::nothrow) b200lc_cuhd_session();
    if (!s) return B200LC_ERR_ARG;
    cudaGetDevice(&s->device);
    s->max_symbols = max_symbols;
    s->max_units = (max_symbols * 13 + 31) / 32 + 2;  // worst case for 13-bit codes + pad unit
    size_t sa = b200lc_cuhd_decode_scratch_bytes(s->max_units);
    size_t sb = b200lc_cuhd_encode_scratch_bytes(max_symbols);
    s->scratch_bytes = sa > sb ? sa : sb;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_symbols, max_symbols + 16);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_units, s->max_units * sizeof(u32));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_scratch, s->scratch_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_small, kSmallBytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_piece_hist, b200lc_cuhd_piece_hist_bytes(max_symbols) + 1024);
    if (e == cudaSuccess) e = cudaHostAlloc(&s->h_small, (257 + kMaxChunks) * sizeof(u64), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        fprintf(stderr, "b200lc: cuhd session allocation failed: %s\n", cudaGetErrorString(e));
        b200lc_cuhd_session_destroy(s);
        return B200LC_ERR_CUDA;
    }
    *out = s;
    return B200LC_OK;
}

extern "C" int b200lc_cuhd_session_destroy(b200lc_cuhd_session *s)
{
    if (!s) return B200LC_OK;
    DeviceScope on(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_symbols);
    cudaFree(s->d_units);
    cudaFree(s->d_scratch);
    cudaFree(s->d_small);
    cudaFree(s->d_piece_hist);
    if (s->h_small) cudaFreeHost(s->h_small);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->h2d) cudaStreamDestroy(s->h2d);
    if (s->d2h) cudaStreamDestroy(s->d2h);
    for (cudaEvent_t ev : s->events) cudaEventDestroy(ev);
    delete s;
    return B200LC_OK;
}

// host symbols -> host stream units + dictionary + LUT.  Synchronous.  Pipelined like the decode
// below: the symbols go up in chunks on one copy engine while the per-piece histograms of the
// chunks that have arrived are computed; the dictionary is built on the host from their sum; the
// one-pass packer (b200lc_cuhd_encode_planned) needs no counting pass because the piece histograms
// and the code lengths give the bit offset of every piece.
extern "C" int b200lc_cuhd_session_encode(b200lc_cuhd_session *s, const uint8_t *h_in, size_t n,
                                          int max_len, uint32_t *h_units, size_t units_cap,
                                          size_t *n_units, uint32_t *h_code_of_symbol,
                                          uint8_t *h_len_of_symbol, uint8_t *h_lut)
{
    if (!s || !h_in || !h_units || !n_units || !h_code_of_symbol || !h_len_of_symbol)
        return B200LC_ERR_ARG;
    if (n == 0 || n > s->max_symbols) return B200LC_ERR_ARG;
    if (max_len < 1 || max_len > 13) return B200LC_ERR_UNSUPPORTED;
    DeviceScope on(s->device);
    u64 *d_hist = reinterpret_cast<u64 *>(s->d_small);
    u32 *d_code = reinterpret_cast<u32 *>(s->d_small + 2048);
    u8 *d_len = s->d_small + 3072;
    u64 *d_bits = reinterpret_cast<u64 *>(s->d_small + 4096 + (size_t(2) << 13));

    const size_t ps = b200lc_cuhd_piece_symbols();
    size_t chunk = std::max<size_t>(1, chunk_bytes(n, size_t(16) << 20, size_t(64) << 20, true) / ps) *
                   ps;                                         // a whole number of pieces
    while ((n + chunk - 1) / chunk > kMaxChunks) chunk *= 2;
    const size_t nchunks = (n + chunk - 1) / chunk;
    while (s->events.size() < 2 * nchunks) {
        cudaEvent_t ev;
        B200LC_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        s->events.push_back(ev);
    }
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t lo = k * chunk, hi = lo + chunk < n ? lo + chunk : n;
        B200LC_CUDA_TRY(cudaMemcpyAsync(s->d_symbols + lo, h_in + lo, hi - lo, cudaMemcpyHostToDevice, s->h2d));
        B200LC_CUDA_TRY(cudaEventRecord(s->events[k], s->h2d));
        B200LC_CUDA_TRY(cudaStreamWaitEvent(s->stream, s->events[k], 0));
        const int rc = b200lc_histogram_u8_pieces_part(s->d_symbols, n, lo / ps, (hi + ps - 1) / ps,
                                                       s->d_piece_hist, s->stream);
        if (rc) { cudaDeviceSynchronize(); return rc; }
    }
    int rc = b200lc_histogram_u8_pieces_finish(s->d_piece_hist, n, d_hist, s->stream);
    if (rc) { cudaDeviceSynchronize(); return rc; }
    B200LC_CUDA_TRY(cudaMemcpyAsync(s->h_small, d_hist, 256 * sizeof(u64), cudaMemcpyDeviceToHost,
                                    s->stream));
    B200LC_CUDA_TRY(cudaStreamSynchronize(s->stream));
    rc = b200lc_cuhd_build_table(s->h_small, max_len, h_code_of_symbol, h_len_of_symbol, h_lut);
    if (rc) return rc;
    const size_t units = b200lc_cuhd_compressed_units(s->h_small, h_len_of_symbol);
    *n_units = units;
    if (units + 1 > units_cap || units + 1 > s->max_units) return B200LC_ERR_OVERFLOW;
    B200LC_CUDA_TRY(cudaMemcpyAsync(d_code, h_code_of_symbol, 256 * sizeof(u32),
                                    cudaMemcpyHostToDevice, s->stream));
    B200LC_CUDA_TRY(cudaMemcpyAsync(d_len, h_len_of_symbol, 256, cudaMemcpyHostToDevice, s->stream));
    rc = b200lc_cuhd_encode_planned(s->d_symbols, n, d_code, d_len, s->d_piece_hist, s->d_units, s->max_units,
                                    d_bits, s->d_scratch, s->scratch_bytes, s->stream);
    if (rc) return rc;
    B200LC_CUDA_TRY(cudaMemcpyAsync(h_units, s->d_units, (units + 1) * sizeof(u32),
                                    cudaMemcpyDeviceToHost, s->stream));
    B200LC_CUDA_TRY(cudaMemcpyAsync(&s->h_small[256], d_bits, sizeof(u64), cudaMemcpyDeviceToHost,
                                    s->stream));
    B200LC_CUDA_TRY(cudaStreamSynchronize(s->stream));
    if ((s->h_small[256] + 31) / 32 != units) return B200LC_ERR_OVERFLOW;
    return B200LC_OK;
}

// host stream units + LUT -> host symbols.  Synchronous.  This is demo.cc:138-168 in one call,
// pipelined: the stream goes up in chunks on one copy engine, pieces are decoded as soon as their
// units (+ 4 lookahead units) have arrived, and the symbols that are final go down on the other
// copy engine while later chunks are still travelling up.
extern "C" int b200lc_cuhd_session_decode(b200lc_cuhd_session *s, const uint32_t *h_units,
                                          size_t n_units, const void *h_lut, int max_len,
                                          uint8_t *h_out, size_t n_out)
{
    if (!s || !h_units || !h_lut || !h_out) return B200LC_ERR_ARG;
    if (n_units == 0 || n_units > s->max_units || n_out > s->max_symbols) return B200LC_ERR_ARG;
    if (max_len < 1 || max_len > 13) return B200LC_ERR_UNSUPPORTED;
    DeviceScope on(s->device);
    u8 *d_lut = s->d_small + 4096;
    B200LC_CUDA_TRY(cudaMemcpyAsync(d_lut, h_lut, size_t(2) << max_len, cudaMemcpyHostToDevice,
                                    s->stream));
    const size_t pu = b200lc_cuhd_decode_piece_units();
    const size_t pieces = (n_units + pu - 1) / pu;
    size_t chunk_units = chunk_bytes(n_units * 4, size_t(8) << 20, size_t(32) << 20, false) / 4;
    while ((n_units + chunk_units - 1) / chunk_units > kMaxChunks) chunk_units *= 2;
    const size_t nchunks = (n_units + chunk_units - 1) / chunk_units;
    while (s->events.size() < 2 * nchunks) {
        cudaEvent_t ev;
        B200LC_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        s->events.push_back(ev);
    }
    u64 *h_prog = s->h_small + 257;
    std::vector<int> decoded(nchunks, 0);
    size_t done_pieces = 0;
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t lo = k * chunk_units, hi = lo + chunk_units < n_units ? lo + chunk_units : n_units;
        B200LC_CUDA_TRY(cudaMemcpyAsync(s->d_units + lo, h_units + lo, (hi - lo) * sizeof(u32),
                                        cudaMemcpyHostToDevice, s->h2d));
        B200LC_CUDA_TRY(cudaEventRecord(s->events[2 * k], s->h2d));
        const size_t end_piece = hi == n_units ? pieces : (hi >= 4 ? (hi - 4) / pu : 0);
        if (end_piece <= done_pieces) continue;
        B200LC_CUDA_TRY(cudaStreamWaitEvent(s->stream, s->events[2 * k], 0));
        int rc = b200lc_cuhd_decode_pieces(s->d_units, n_units, s->d_symbols, n_out, d_lut, max_len,
                                           s->d_scratch, s->scratch_bytes, done_pieces, end_piece,
                                           s->stream);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        rc = b200lc_cuhd_decode_progress_async(s->d_scratch, end_piece, &h_prog[k], s->stream);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        B200LC_CUDA_TRY(cudaEventRecord(s->events[2 * k + 1], s->stream));
        decoded[k] = 1;
        done_pieces = end_piece;
    }
    size_t copied = 0;
    for (size_t k = 0; k < nchunks; ++k) {
        if (!decoded[k]) continue;
        B200LC_CUDA_TRY(cudaEventSynchronize(s->events[2 * k + 1]));
        size_t ready = (size_t)(h_prog[k] & ((u64(1) << 56) - 1));
        if (ready > n_out || k + 1 == nchunks) ready = n_out;
        if (ready > copied) {
            B200LC_CUDA_TRY(cudaMemcpyAsync(h_out + copied, s->d_symbols + copied, ready - copied,
                                            cudaMemcpyDeviceToHost, s->d2h));
            copied = ready;
        }
    }
    B200LC_CUDA_TRY(cudaStreamSynchronize(s->d2h));
    B200LC_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return B200LC_OK;
}
