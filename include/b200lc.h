/*
 * b200lc.h -- C ABI of libb200lc.so, the Blackwell (sm_100a) kernel suite behind the entry points
 * of the compressors vendored in dingwentao/GPU-lossless-compression.
 *
 * Two kinds of symbols are exported:
 *   1. b200lc_*  : explicit-stream, explicit-scratch, device-pointer entry points.  No hidden
 *                  allocation, no host synchronisation unless stated.  These are what the
 *                  reference-named shims below are built on and what bench.py times.
 *   2. the reference's own names (include/cudpp.h, include/culzss_gpu.h, include/cuhd_c.h):
 *                  same names, argument meaning and error behaviour as the reference functions
 *                  they replace, so a reference caller links against libb200lc.so unchanged.
 *
 * All reference citations are relative to /root/reference (dingwentao/GPU-lossless-compression).
 * Every function returns B200LC_OK (0) or a negative B200LC_ERR_* code unless stated otherwise.
 * There is no CPU fallback anywhere: a missing GPU is an error.
 */
#ifndef B200LC_H_
#define B200LC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LC_OK 0
#define B200LC_ERR_ARG (-1)         /* null / misaligned pointer, bad size */
#define B200LC_ERR_CUDA (-2)        /* a CUDA runtime call failed (message on stderr) */
#define B200LC_ERR_SCRATCH (-3)     /* scratch buffer smaller than *_scratch_bytes() */
#define B200LC_ERR_UNSUPPORTED (-4) /* parameter outside the supported range */
#define B200LC_ERR_OVERFLOW (-5)    /* output capacity exceeded */

/* Version / build identification: "b200lc <git-describe or 'dev'> sm_100a". */
const char *b200lc_version(void);

/* ------------------------------------------------------------------------------------------
 * Hot path 3: canonical-Huffman decode of a CUHD stream.
 * Replaces cuhd::CUHDGPUDecoder::decode (cuhd-icpp/src/cuhd_gpu_decoder.cu:422-523,
 * declared cuhd-icpp/include/cuhd_gpu_decoder.h:24-32).
 *
 *   d_units      compressed stream, uint32 units, codes MSB-first inside each unit
 *                (cuhd-icpp/encoder/src/llhuffman_encoder.cc:200-238); may or may not include
 *                the reference's trailing zero pad unit (cuhd_input_buffer.cc:20-27): units past
 *                n_units read as zero.
 *   n_units      number of units readable at d_units
 *   d_out        n_out decoded symbols (uint8)
 *   d_table      flat LUT of (1 << max_codeword_length) entries {uint8 num_bits, uint8 symbol}
 *                = cuhd::CUHDCodetableItemSingle (cuhd-icpp/include/cuhd_codetable.h:20-23)
 *   max_codeword_length   11 in the reference (cuhd_constants.h:15); 1..13 accepted
 *   d_scratch    >= b200lc_cuhd_decode_scratch_bytes(n_units) bytes, 128-byte aligned; replaces
 *                CUHDGPUDecoderMemory (cuhd_gpu_decoder_memory.cc:25-48; 20 B per 16 B of input)
 *                by 128 B per 4 KiB of input
 *   stream       cudaStream_t (NULL = default stream).  Asynchronous: returns after enqueue.
 * All pointers are device pointers.
 */
size_t b200lc_cuhd_decode_scratch_bytes(size_t n_units);
int b200lc_cuhd_decode(const uint32_t *d_units, size_t n_units, uint8_t *d_out, size_t n_out,
                       const void *d_table, int max_codeword_length, void *d_scratch,
                       size_t scratch_bytes, void *stream);

/* Piece-wise decoding for pipelines that overlap the H2D copy of the stream with decoding and the
 * D2H copy of the symbols (what b200lc_cuhd_session_decode does).  A piece is
 * b200lc_cuhd_decode_piece_units() units; pieces must be decoded in increasing order with the
 * same arguments and scratch; decoding pieces [first, end) needs the units of those pieces plus
 * 4 to be resident (or end to be the last piece).  b200lc_cuhd_decode_progress_async copies to
 * *h_symbols (pinned) a word whose low 56 bits = number of leading output symbols that are final
 * once pieces [0, end_piece) are done. */
size_t b200lc_cuhd_decode_piece_units(void);
int b200lc_cuhd_decode_pieces(const uint32_t *d_units, size_t n_units, uint8_t *d_out, size_t n_out,
                              const void *d_table, int max_codeword_length, void *d_scratch,
                              size_t scratch_bytes, size_t first_piece, size_t end_piece,
                              void *stream);
int b200lc_cuhd_decode_progress_async(const void *d_scratch, size_t end_piece,
                                      uint64_t *h_symbols, void *stream);

/* Batch: many independent streams that share ONE code table (e.g. the blocks of a buffer packed
 * separately for random access), decoded by one launch.  Stream i occupies
 * d_units[unit_offset .. unit_offset + n_units) and decodes to d_out[out_offset .. out_offset + n_out);
 * its codes start at bit 0 of its first unit.  unit_offset % 4 == 0 lets the stream use TMA bulk
 * copies (any offset works).  The stream descriptors are HOST memory; everything else as above.
 * Scratch: b200lc_cuhd_decode_batch_scratch_bytes(), 128-byte aligned.  Asynchronous once the
 * descriptors have been staged. */
typedef struct b200lc_cuhd_stream {
    uint64_t unit_offset, n_units, out_offset, n_out;
} b200lc_cuhd_stream;
size_t b200lc_cuhd_decode_batch_scratch_bytes(const b200lc_cuhd_stream *h_streams, size_t n_streams);
int b200lc_cuhd_decode_batch(const uint32_t *d_units, uint8_t *d_out, const b200lc_cuhd_stream *h_streams,
                             size_t n_streams, const void *d_table, int max_codeword_length,
                             void *d_scratch, size_t scratch_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * CUHD-format encoder side: replaces the sequential CPU stages of the
 * reference's demo (cuhd-icpp/src/demo.cc:90-107).
 */

/* 256-bin histogram of n bytes (replaces the frequency count at llhuffman_encoder.cc:23-26).
 * d_in must be 16-byte aligned; d_hist[256] is overwritten.  Asynchronous. */
int b200lc_histogram_u8(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *stream);

/* HOST function: optimal length-limited prefix code for a 256-bin histogram, canonical codes and
 * flat decode LUT.  Counterpart of get_symbol_lengths / get_encoder_table / get_decoder_table
 * (llhuffman_encoder.cc:18-198,240-262), with exact integer weights and (length, symbol) tie
 * order.  All pointers are host pointers; lut (may be NULL) receives (1 << max_len) entries of
 * {uint8 num_bits, uint8 symbol}; len_of_symbol[s] == 0 means "symbol absent". */
int b200lc_cuhd_build_table(const uint64_t *hist, int max_len, uint32_t *code_of_symbol,
                            uint8_t *len_of_symbol, uint8_t *lut);
size_t b200lc_cuhd_compressed_units(const uint64_t *hist, const uint8_t *len_of_symbol);

/* MSB-first bit packer, replaces LLHuffmanEncoder::encode_memory (llhuffman_encoder.cc:200-238).
 *   d_in[n]            symbols
 *   d_code_of_symbol   uint32[256] right-aligned codewords, d_len_of_symbol uint8[256] lengths
 *   d_units            output, 16-byte aligned, capacity units_cap; ceil(bits/32) units are
 *                      written, plus one zero pad unit (cuhd_input_buffer.cc:20-27) if it fits
 *   d_total_bits       device uint64 receiving the number of stream bits
 *   d_scratch          >= b200lc_cuhd_encode_scratch_bytes(n), 128-byte aligned
 * Asynchronous.  b200lc_cuhd_encode_overflowed() synchronises the stream and reports
 * B200LC_ERR_OVERFLOW if units_cap was too small for the last call on that scratch. */
size_t b200lc_cuhd_encode_scratch_bytes(size_t n);
int b200lc_cuhd_encode(const uint8_t *d_in, size_t n, const uint32_t *d_code_of_symbol,
                       const uint8_t *d_len_of_symbol, uint32_t *d_units, size_t units_cap,
                       uint64_t *d_total_bits, void *d_scratch, size_t scratch_bytes,
                       void *stream);
int b200lc_cuhd_encode_overflowed(const void *d_scratch, void *stream);

/* One-pass variant for callers that run the histogram anyway: b200lc_histogram_u8_pieces also keeps
 * one 256-bin histogram per 128 KiB piece (d_piece_hist, b200lc_cuhd_piece_hist_bytes(n) bytes);
 * with them and the code lengths b200lc_cuhd_encode_planned knows the bit offset of every piece
 * before it starts, so the packer's counting pass and its look-back disappear.  Same arguments,
 * same output and same overflow reporting as b200lc_cuhd_encode.  Asynchronous. */
size_t b200lc_cuhd_piece_hist_bytes(size_t n);
int b200lc_histogram_u8_pieces(const uint8_t *d_in, size_t n, uint64_t *d_hist, uint32_t *d_piece_hist,
                               void *stream);
/* The same in two steps for input that arrives in chunks: the histograms of pieces
 * [first_piece, end_piece) (a piece = b200lc_cuhd_piece_symbols() symbols of d_in[n]) as soon as
 * their symbols are resident, then the reduction of all pieces into d_hist[256]. */
size_t b200lc_cuhd_piece_symbols(void);
int b200lc_histogram_u8_pieces_part(const uint8_t *d_in, size_t n, size_t first_piece, size_t end_piece,
                                    uint32_t *d_piece_hist, void *stream);
int b200lc_histogram_u8_pieces_finish(const uint32_t *d_piece_hist, size_t n, uint64_t *d_hist, void *stream);
int b200lc_cuhd_encode_planned(const uint8_t *d_in, size_t n, const uint32_t *d_code_of_symbol,
                               const uint8_t *d_len_of_symbol, const uint32_t *d_piece_hist,
                               uint32_t *d_units, size_t units_cap, uint64_t *d_total_bits,
                               void *d_scratch, size_t scratch_bytes, void *stream);

/* Blocks: d_in[n] is cut into blocks of block_symbols symbols (the last one may be shorter) and
 * every block is packed as an independent stream with the same dictionary, all by one launch:
 * block b starts at bit 0 of d_units + b * unit_stride (unit_stride a multiple of 4, at least
 * the block's units + 1 for the pad unit), d_block_bits[b] receives its number of stream bits.
 * These are the streams b200lc_cuhd_decode_batch takes (unit_offset = b * unit_stride,
 * n_units = ceil(bits / 32)).  block_symbols % 16 == 0 lets the blocks use TMA bulk copies.
 * Asynchronous; b200lc_cuhd_encode_overflowed() reports a unit_stride that was too small. */
size_t b200lc_cuhd_encode_blocks_scratch_bytes(size_t n, size_t block_symbols);
int b200lc_cuhd_encode_blocks(const uint8_t *d_in, size_t n, size_t block_symbols,
                              const uint32_t *d_code_of_symbol, const uint8_t *d_len_of_symbol,
                              uint32_t *d_units, size_t unit_stride, uint64_t *d_block_bits,
                              void *d_scratch, size_t scratch_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer session for the CUHD path: the reference demo's flow around the decoder
 * (cuhd-icpp/src/demo.cc:122-168: device buffers, H2D of table + stream, decode, D2H) and around
 * the CPU encoder (demo.cc:90-107) as two synchronous calls.  The session owns device buffers
 * for up to max_symbols symbols, the scratch and a stream.  Host pointers should be pinned.
 */
typedef struct b200lc_cuhd_session b200lc_cuhd_session;
int b200lc_cuhd_session_create(size_t max_symbols, b200lc_cuhd_session **out);
int b200lc_cuhd_session_destroy(b200lc_cuhd_session *s);
/* h_in[n] -> h_units[*n_units (+1 pad unit)], dictionary (h_code_of_symbol[256],
 * h_len_of_symbol[256]) and LUT (h_lut, (1 << max_len) * 2 bytes, may be NULL). */
int b200lc_cuhd_session_encode(b200lc_cuhd_session *s, const uint8_t *h_in, size_t n, int max_len,
                               uint32_t *h_units, size_t units_cap, size_t *n_units,
                               uint32_t *h_code_of_symbol, uint8_t *h_len_of_symbol,
                               uint8_t *h_lut);
/* h_units[n_units] + h_lut -> h_out[n_out]. */
int b200lc_cuhd_session_decode(b200lc_cuhd_session *s, const uint32_t *h_units, size_t n_units,
                               const void *h_lut, int max_len, uint8_t *h_out, size_t n_out);

/* ------------------------------------------------------------------------------------------
 * Hot path 2: CULZSS-compatible LZSS (cuda-lzss-cluster), device-pointer batch API.
 * Bit-exact with the reference's EncodeKernel + aftercomp + trailer
 * (gpu_compress.cu:104-350, :462-673) and DecodeKernel (gpu_decompress.cu:120-244):
 * WINDOW_SIZE 128, MAX_CODED 128, 4096-byte packets, per-buffer trailer
 * [packets][npk x u16 BE sizes][u32 BE buf_length][u16 BE pad = 0].
 *
 * Encode: nbuf independent buffers of buf_length bytes each (multiple of 4096; the reference
 * uses 1 MiB, main.c:62), contiguous at d_in (16-byte aligned).  Buffer b is written to
 * d_out + b * out_stride and its size incl. trailer to d_comp_len[b]; d_comp_len[b] == 0 means
 * the reference would have reported "compression took more" (aftercompression_wrapper returns
 * 0, gpu_compress.cu:494-498,661-662) and the caller stores the buffer raw (culzss.c:177-183).
 * out_stride >= buf_length + buf_length / 8 + 1024 is always sufficient.  Asynchronous.
 */
size_t b200lc_culzss_encode_scratch_bytes(size_t nbuf, size_t buf_length);
int b200lc_culzss_encode_batch(const uint8_t *d_in, size_t nbuf, size_t buf_length, uint8_t *d_out,
                               size_t out_stride, uint32_t *d_comp_len, void *d_scratch,
                               size_t scratch_bytes, void *stream);
/* The same with the kernel chosen by the caller.  Both kernels write the reference's bytes:
 * B200LC_CULZSS_KERNEL_CTA   one CTA per packet, a match for every position, selection and packing
 *                            afterwards (any batch size);
 * B200LC_CULZSS_KERNEL_LANE  one packet per GPU lane, the reference's match finder evaluated only at
 *                            the positions the greedy selection visits (csrc/culzss_lane.cuh); 3-4x
 *                            the throughput once >= ~10^5 packets are in flight;
 * B200LC_CULZSS_KERNEL_AUTO  what b200lc_culzss_encode_batch does: CTA below 40960 packets (160 MiB) per
 *                            call; from there on the CTA kernel codes a probe of 512 packets, a
 *                            one-CTA kernel looks at the probe's ratio and both kernels are launched
 *                            on the rest, one of which returns at once -- LANE if the probe shrank
 *                            to half or less (frequent matches), else CTA.  No host round trip.
 *                            Environment B200LC_CULZSS_PARITY_LANE=0|1 overrides (never / always). */
#define B200LC_CULZSS_KERNEL_AUTO 0
#define B200LC_CULZSS_KERNEL_CTA 1
#define B200LC_CULZSS_KERNEL_LANE 2
int b200lc_culzss_encode_batch_ex(const uint8_t *d_in, size_t nbuf, size_t buf_length, uint8_t *d_out,
                                  size_t out_stride, uint32_t *d_comp_len, void *d_scratch,
                                  size_t scratch_bytes, int kernel, void *stream);
/* FAST MODE -- NOT bit-exact with the reference encoder (every table and test labels it
 * non-parity).  Same buffer format, token format, window (128) and packets (4096) as
 * b200lc_culzss_encode_batch, so the output decodes with the reference's DecodeKernel
 * (gpu_decompress.cu:164-242) and with b200lc_culzss_decode_batch, but the match finder is a
 * shared-memory hash chain over three-byte prefixes that looks at the `depth` (1, 2 or 4) most
 * recent candidates instead of the reference's exhaustive streak scanner: ~20-35x the throughput of
 * parity mode for a 10-25 % larger output.  Same arguments otherwise.
 *
 * depth = B200LC_CULZSS_FAST_LANE selects the second fast formulation (also NON-PARITY, same
 * format): one packet per GPU lane, a greedy parse through a lane-private 128-entry hash of
 * three-byte prefixes with tokens and flag bytes emitted in the same serial walk
 * (csrc/culzss_lane.cuh).  Matches are at most 108 bytes long.  It needs >= ~10^5 packets in flight
 * to fill the GPU (one packet per lane), i.e. it is the mode for large batches. */
#define B200LC_CULZSS_FAST_LANE (-1)
int b200lc_culzss_encode_fast_batch(const uint8_t *d_in, size_t nbuf, size_t buf_length, uint8_t *d_out,
                                    size_t out_stride, uint32_t *d_comp_len, void *d_scratch,
                                    size_t scratch_bytes, int depth, void *stream);
/* Decode: compressed buffer b occupies d_comp[d_comp_offsets[b] .. d_comp_offsets[b+1]) and is
 * decoded to d_out + b * buf_length (d_out 16-byte aligned).  A buffer whose stored size equals
 * buf_length is raw and copied (decompression.c:90-108, deculzss.c:94-95).  Asynchronous. */
size_t b200lc_culzss_decode_scratch_bytes(size_t nbuf, size_t buf_length);
int b200lc_culzss_decode_batch(const uint8_t *d_comp, const uint64_t *d_comp_offsets, size_t nbuf,
                               size_t buf_length, uint8_t *d_out, void *d_scratch,
                               size_t scratch_bytes, void *stream);

/* CULZSS file container, HOST buffers, synchronous: the byte layout the
 * reference CLI writes and reads (cuda-lzss-cluster/main.c:236-245, culzss.c:220,243-264,
 * decompression.c:90-141): u32 nblocks, u32 padding, u32 cumulative_end[nblocks], buffers of
 * 1 MiB input each (raw when the stored size is 1 MiB).  n >= 1 MiB like the reference
 * (main.c:225-229).  Byte-identical to the reference's file for n % 1 MiB == 0; otherwise the
 * last buffer is zero-padded where the reference leaves stale bytes (main.c:122-130). */
size_t b200lc_culzss_container_bound(size_t n);
int b200lc_culzss_compress_container(const uint8_t *h_in, size_t n, uint8_t *h_out, size_t cap,
                                     size_t *out_len);
int b200lc_culzss_decompress_container(const uint8_t *h_in, size_t n, uint8_t *h_out, size_t cap,
                                       size_t *out_len);

/* ------------------------------------------------------------------------------------------
 * Hot path 1: BWT -> MTF -> Huffman as in cudppCompress (cudpp-inpar), batched over independent
 * blocks.  The CUDPP-named single-block entry points are in include/cudpp.h.
 * All *_batch functions take nblocks blocks of n bytes, contiguous at stride n.
 */

/* Suffix-array BWT (cudppBurrowsWheelerTransform, compress_app.cu:243-267 + sa_app.cu:125-391):
 * d_out[b*n + i] = last column, d_index[b] = row of the original string.  n < 2^21,
 * nblocks * n < 2^30 per call.  Synchronises the stream (one counter read per doubling round). */
size_t b200lc_bwt_scratch_bytes(size_t nblocks, size_t n);
int b200lc_bwt_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint8_t *d_out, int *d_index,
                     void *d_scratch, size_t scratch_bytes, void *stream);
/* cudppSuffixArray: d_sa[b*n + j] = start (inside block b) of its j-th smallest suffix. */
int b200lc_suffix_array_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint32_t *d_sa,
                              void *d_scratch, size_t scratch_bytes, void *stream);

/* Move-to-front with initial list 0..255 (cudppMoveToFrontTransform, compress_app.cu:133-223;
 * gold test_compress.cpp:93-125).  Any n.  Asynchronous. */
size_t b200lc_mtf_scratch_bytes(size_t nblocks, size_t n);
int b200lc_mtf_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint8_t *d_out, void *d_scratch,
                     size_t scratch_bytes, void *stream);

/* Huffman stage of cudppCompress (compress_app.cu:65-117) on MTF output.  Per block b:
 * d_hist[b*256..] histogram, d_offsets[b*nhb..] word offset of each 4096-symbol block's [nWords]
 * cell (nhb = ceil(n/4096)), d_total_words[b], stream at d_out + b*out_stride_words.
 * *d_error (device uint32) != 0: a code is longer than 32 bits, a 4096-symbol block needs more
 * than 1536 words (the reference's hard capacity) or out_stride_words is too small.
 * Asynchronous. */
size_t b200lc_cudpp_huffman_scratch_bytes(size_t nblocks, size_t n);
int b200lc_cudpp_huffman_batch(const uint8_t *d_mtf, size_t nblocks, size_t n, uint32_t *d_hist,
                               uint32_t *d_offsets, uint32_t *d_total_words, uint32_t *d_out,
                               size_t out_stride_words, uint32_t *d_error, void *d_scratch,
                               size_t scratch_bytes, void *stream);

/* cudppCompress for a batch of blocks (compress_app.cu:507-526). */
size_t b200lc_cudpp_compress_scratch_bytes(size_t nblocks, size_t n);
int b200lc_cudpp_compress_batch(const uint8_t *d_in, size_t nblocks, size_t n, int *d_bwt_index,
                                uint32_t *d_hist, uint32_t *d_offsets, uint32_t *d_total_words,
                                uint32_t *d_out, size_t out_stride_words, uint32_t *d_error,
                                void *d_scratch, size_t scratch_bytes, void *stream);

/* ---- decoder for the cudppCompress stream -----------------------------
 * The reference ships no GPU decoder; its CPU gold is computeCompressGold
 * (cudpp-inpar/apps/cudpp_testrig/test_compress.cpp:192-364).  These entry points invert
 * b200lc_cudpp_compress_batch / cudppCompress stage by stage, batched over blocks.
 * Limits: n < 2^24, nblocks * n < 2^30.  *d_error != 0 after the stream has drained = corrupt
 * input (4 offsets outside the stream, 5 invalid code / block too short, 6 bwt index >= n). */
size_t b200lc_inverse_mtf_scratch_bytes(size_t nblocks, size_t n);
int b200lc_inverse_mtf_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint8_t *d_out,
                             void *d_scratch, size_t scratch_bytes, void *stream);
size_t b200lc_inverse_bwt_scratch_bytes(size_t nblocks, size_t n);
int b200lc_inverse_bwt_batch(const uint8_t *d_bwt, const int *d_bwt_index, size_t nblocks, size_t n,
                             uint8_t *d_out, uint32_t *d_error, void *d_scratch,
                             size_t scratch_bytes, void *stream);
/* The same for ONE block whose alphabet has no end marker (libbsc's bsc_bwt_encode output, bwt.h:38-61):
 * d_u[0] = T[n-1] followed by the last column without the row of suffix 0, primary = that row + 1
 * (1..n).  The end marker is virtual (row 0 of an (n + 1)-row problem).  d_out needs n + 1 bytes,
 * the first n are the block.  Blocks of 2^24 rows and more use 64-bit row entries.  Synchronous
 * at its end (it copies `primary` to the device). */
size_t b200lc_inverse_bwt_primary_scratch_bytes(size_t n);
int b200lc_inverse_bwt_primary(const uint8_t *d_u, size_t n, int primary, uint8_t *d_out, uint32_t *d_error,
                               void *d_scratch, size_t scratch_bytes, void *stream);

size_t b200lc_cudpp_decompress_scratch_bytes(size_t nblocks, size_t n);
int b200lc_cudpp_decompress_batch(const int *d_bwt_index, const uint32_t *d_hist,
                                  const uint32_t *d_offsets, const uint32_t *d_comp,
                                  size_t comp_stride_words, size_t nblocks, size_t n,
                                  uint8_t *d_out, uint32_t *d_error, void *d_scratch,
                                  size_t scratch_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Device-wide primitives under the BWT paths (csrc/devprims.cu): hand-written replacements for
 * the library sorts/scans the reference leans on (CUB sorts + moderngpu merge in
 * cudpp-inpar/src/cudpp/app/sa_app.cu:125-298, Thrust sorts in
 * cuda-bzip2-ipdpsw/gpuBWTSort.cu:290-418, thrust::exclusive_scan in
 * cuhd-icpp/src/cuhd_gpu_decoder.cu:498-509).  Exported so that they can be tested on their own.
 *
 * Stable LSD radix sort of (key, uint32 value) pairs on key bits [begin_bit, end_bit), every
 * segment of seg_len consecutive elements sorted independently (seg_len == 0 or >= n: one
 * segment).  The pairs start in (d_keys_a, d_vals_a) and ping-pong with (d_keys_b, d_vals_b);
 * *result_in_b (host int) tells where they ended up.  n < 2^30.  Asynchronous. */
size_t b200lc_sort_scratch_bytes(size_t n, size_t seg_len);
int b200lc_sort_pairs_u64(uint64_t *d_keys_a, uint64_t *d_keys_b, uint32_t *d_vals_a,
                          uint32_t *d_vals_b, size_t n, size_t seg_len, int begin_bit, int end_bit,
                          void *d_scratch, size_t scratch_bytes, void *stream, int *result_in_b);
int b200lc_sort_pairs_u32(uint32_t *d_keys_a, uint32_t *d_keys_b, uint32_t *d_vals_a,
                          uint32_t *d_vals_b, size_t n, size_t seg_len, int begin_bit, int end_bit,
                          void *d_scratch, size_t scratch_bytes, void *stream, int *result_in_b);
/* Single-pass scans over n uint32 (d_in == d_out allowed): exclusive sum, inclusive running max. */
size_t b200lc_scan_scratch_bytes(size_t n);
int b200lc_exclusive_sum_u32(const uint32_t *d_in, uint32_t *d_out, size_t n, void *d_scratch,
                             size_t scratch_bytes, void *stream);
int b200lc_inclusive_max_u32(const uint32_t *d_in, uint32_t *d_out, size_t n, void *d_scratch,
                             size_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_H_ */
