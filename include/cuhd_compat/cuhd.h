/*
 * cuhd.h (b200lc) -- source-compatible mirror of the CUHD host interface.
 *
 * cuhd-icpp's boundary is a C++ one: cuhd::CUHDGPUDecoder::decode (static) on RAII buffer
 * classes, shipped as lib/cuhd.a (cuhd-icpp/include/cuhd.h:11-20, Makefile:38-39).  This
 * header-only mirror keeps the class names, constructors and member functions the reference's
 * demo uses (cuhd-icpp/src/demo.cc:118-178), implemented on the C ABI of libb200lc.so, so
 * demo.cc compiles UNCHANGED with   -I <repo>/include/cuhd_compat -I <repo>/include
 * and links with                     -L <repo>/gpu-lossless-compression_b200/lib -lb200lc -lcudart
 *
 * Reference interfaces mirrored (file:line in cuhd-icpp/):
 *   constants / macros      include/cuhd_constants.h:15-24, cuhd_util.h:29-40, cuhd_cuda_definitions.h:21-28
 *   CUHDCodetable           include/cuhd_codetable.h:20-44
 *   CUHDInputBuffer         include/cuhd_input_buffer.h:19-45, src/cuhd_input_buffer.cc:13-32 (+1 zero pad unit)
 *   CUHDOutputBuffer        include/cuhd_output_buffer.h:19-41
 *   CUHDGPUMemoryBuffer<T>  include/cuhd_gpu_memory_buffer.h:18-40 (+ the three subclasses)
 *   CUHDGPUDecoderMemory    include/cuhd_gpu_decoder_memory.h:20-47 (here: the decoder's scratch)
 *   CUHDGPUDecoder::decode  include/cuhd_gpu_decoder.h:24-32
 *   CUHDUtil                include/cuhd_util.h:20-45
 */
#ifndef B200LC_CUHD_COMPAT_H_
#define B200LC_CUHD_COMPAT_H_

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <utility>

#include <cuda_runtime.h>
#include <cuda_runtime_api.h>

#include "b200lc.h"

#define MAX_CODEWORD_LENGTH 11
#define UNIT_TYPE std::uint32_t
#define SYMBOL_TYPE std::uint8_t
#define BIT_COUNT_TYPE std::uint8_t

#define cuhd_buf(TYPE, IDENTIFIER) std::unique_ptr<TYPE[]> IDENTIFIER
#define cuhd_out_buf(TYPE, IDENTIFIER) std::unique_ptr<TYPE[]> IDENTIFIER

#define CUERR {                                                              \
    cudaError_t err;                                                         \
    if ((err = cudaGetLastError()) != cudaSuccess) {                         \
       std::cout << "CUDA error: " << cudaGetErrorString(err) << " : "       \
                 << __FILE__ << ", line " << __LINE__ << std::endl;          \
       exit(1);                                                              \
    }                                                                        \
}

#define SDIV(n, m) ((n + m - 1) / m)
#define TIMER_START(vec, label) vec.push_back(cuhd::CUHDUtil::time(label, [&]() {
#define TIMER_STOP }));

namespace cuhd {

struct CUHDCodetableItemSingle {
    BIT_COUNT_TYPE num_bits;
    SYMBOL_TYPE symbol;
};

class CUHDCodetable {
   public:
    explicit CUHDCodetable(size_t num_entries)
        : size_(size_t(1) << MAX_CODEWORD_LENGTH), num_entries_(num_entries),
          table_(std::make_unique<CUHDCodetableItemSingle[]>(size_)) {}
    size_t get_size() { return size_; }
    size_t get_num_entries() { return num_entries_; }
    size_t get_max_codeword_length() { return MAX_CODEWORD_LENGTH; }
    CUHDCodetableItemSingle* get() { return table_.get(); }

   private:
    size_t size_, num_entries_;
    cuhd_buf(CUHDCodetableItemSingle, table_);
};

class CUHDInputBuffer {
   public:
    CUHDInputBuffer(std::uint8_t* buffer, size_t size) : compressed_size_(size) {
        compressed_size_units_ = (size + sizeof(UNIT_TYPE) - 1) / sizeof(UNIT_TYPE) + 1;  // + pad unit
        buffer_ = std::make_unique<UNIT_TYPE[]>(compressed_size_units_);
        buffer_[compressed_size_units_ - 1] = 0;
        if (compressed_size_units_ > 1) buffer_[compressed_size_units_ - 2] = 0;
        std::copy(buffer, buffer + size, reinterpret_cast<std::uint8_t*>(buffer_.get()));
    }
    UNIT_TYPE* get_compressed_data() { return buffer_.get(); }
    size_t get_compressed_size() { return compressed_size_; }
    size_t get_compressed_size_units() { return compressed_size_units_; }
    size_t get_unit_size() { return sizeof(UNIT_TYPE); }

   private:
    size_t compressed_size_, compressed_size_units_;
    cuhd_buf(UNIT_TYPE, buffer_);
};

class CUHDOutputBuffer {
   public:
    explicit CUHDOutputBuffer(size_t size)
        : uncompressed_size_(size), buffer_(std::make_unique<SYMBOL_TYPE[]>(size)) {}
    std::unique_ptr<SYMBOL_TYPE[]>& get_decompressed_data() { return buffer_; }
    size_t get_uncompressed_size() { return uncompressed_size_; }
    size_t get_symbol_size() { return sizeof(SYMBOL_TYPE); }

   private:
    size_t uncompressed_size_;
    cuhd_out_buf(SYMBOL_TYPE, buffer_);
};

template <typename T>
class CUHDGPUMemoryBuffer {
   public:
    CUHDGPUMemoryBuffer(T* buffer, size_t size)
        : buffer_(buffer), buffer_device_(nullptr), is_allocated_(false), size_(size) {}
    ~CUHDGPUMemoryBuffer() { free(); }
    T* get() { return buffer_device_; }
    void allocate() {
        if (is_allocated_) return;
        cudaMalloc(reinterpret_cast<void**>(&buffer_device_), size_ * sizeof(T) + 16);
        CUERR
        is_allocated_ = true;
    }
    void free() {
        if (!is_allocated_) return;
        cudaFree(buffer_device_);
        is_allocated_ = false;
    }
    void cpy_host_to_device() {
        cudaMemcpy(buffer_device_, buffer_, size_ * sizeof(T), cudaMemcpyHostToDevice);
        CUERR
    }
    void cpy_device_to_host() {
        cudaMemcpy(buffer_, buffer_device_, size_ * sizeof(T), cudaMemcpyDeviceToHost);
        CUERR
    }

   private:
    T* buffer_;
    T* buffer_device_;
    bool is_allocated_;
    size_t size_;
};

class CUHDGPUInputBuffer : public CUHDGPUMemoryBuffer<UNIT_TYPE> {
   public:
    explicit CUHDGPUInputBuffer(std::shared_ptr<CUHDInputBuffer> b)
        : CUHDGPUMemoryBuffer<UNIT_TYPE>(b->get_compressed_data(), b->get_compressed_size_units()),
          input_buffer_(b) {}

   private:
    std::shared_ptr<CUHDInputBuffer> input_buffer_;
};

class CUHDGPUOutputBuffer : public CUHDGPUMemoryBuffer<SYMBOL_TYPE> {
   public:
    explicit CUHDGPUOutputBuffer(std::shared_ptr<CUHDOutputBuffer> b)
        : CUHDGPUMemoryBuffer<SYMBOL_TYPE>(b->get_decompressed_data().get(), b->get_uncompressed_size()),
          output_buffer_(b) {}

   private:
    std::shared_ptr<CUHDOutputBuffer> output_buffer_;
};

class CUHDGPUCodetable : public CUHDGPUMemoryBuffer<CUHDCodetableItemSingle> {
   public:
    explicit CUHDGPUCodetable(std::shared_ptr<CUHDCodetable> t)
        : CUHDGPUMemoryBuffer<CUHDCodetableItemSingle>(t->get(), t->get_size()), table_(t) {}

   private:
    std::shared_ptr<CUHDCodetable> table_;
};

// The reference keeps 20 bytes of sync state per 16 bytes of input here
// (cuhd_gpu_decoder_memory.cc:25-48); this object only owns the decoder's small scratch.
class CUHDGPUDecoderMemory {
   public:
    CUHDGPUDecoderMemory(size_t num_units, size_t /*subsequence_size*/, size_t /*num_threads*/)
        : bytes_(b200lc_cuhd_decode_scratch_bytes(num_units)), scratch_(nullptr) {}
    ~CUHDGPUDecoderMemory() { free(); }
    void allocate() {
        if (scratch_) return;
        cudaMalloc(&scratch_, bytes_);
        CUERR
    }
    void free() {
        if (scratch_) cudaFree(scratch_);
        scratch_ = nullptr;
    }
    void* get_scratch() { return scratch_; }
    size_t get_scratch_bytes() { return bytes_; }

   private:
    size_t bytes_;
    void* scratch_;
};

class CUHDGPUDecoder {
   public:
    // preferred_subsequence_size / threads_per_block: accepted for source compatibility, the
    // kernel chooses its own decomposition.  Asynchronous like the reference (default stream).
    static void decode(std::shared_ptr<CUHDGPUInputBuffer> input, size_t input_size,
                       std::shared_ptr<CUHDGPUOutputBuffer> output, size_t output_size,
                       std::shared_ptr<CUHDGPUCodetable> table,
                       std::shared_ptr<CUHDGPUDecoderMemory> aux, size_t max_codeword_length,
                       size_t /*preferred_subsequence_size*/, size_t /*threads_per_block*/) {
        const int rc = b200lc_cuhd_decode(input->get(), input_size, output->get(), output_size,
                                          table->get(), static_cast<int>(max_codeword_length),
                                          aux->get_scratch(), aux->get_scratch_bytes(), nullptr);
        if (rc != B200LC_OK) {
            std::cout << "b200lc_cuhd_decode failed: " << rc << std::endl;
            exit(1);
        }
        CUERR
    }
};

class CUHDUtil {
   public:
    static std::pair<std::string, size_t> time(std::string s, std::function<void()> f) {
        auto start = std::chrono::high_resolution_clock::now();
        f();
        auto end = std::chrono::high_resolution_clock::now();
        return {s, static_cast<size_t>(
                       std::chrono::duration_cast<std::chrono::microseconds>(end - start).count())};
    }
    static bool equals(SYMBOL_TYPE* a, SYMBOL_TYPE* b, size_t size) {
        for (size_t i = 0; i < size; ++i)
            if (a[i] != b[i]) return false;
        return true;
    }
};

}  // namespace cuhd
#endif /* B200LC_CUHD_COMPAT_H_ */
