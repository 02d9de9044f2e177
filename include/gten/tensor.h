// gten/tensor.h -- gten::Tensor over HBM storage (reference gten/tensor.h:20-137, gten/tensor.cpp:27-201).
//
// Same public surface as the reference: <= 3-D shape/strides over a shared buffer, block-aware byte sizing, resize
// without reallocation, views, raw data_ptr<T>() access.  The difference is where the bytes live: every tensor has a
// device buffer (gtb_malloc) and a lazily created host mirror.  Callers of the reference API write weights and read
// logits through raw host pointers (tinyllama.cpp:320, 414), so coherence is tracked per buffer:
//   data_ptr()   -> host mirror made current (D2H if the device copy is newer), host assumed modified afterwards;
//   device_in()  -> device copy made current (H2D if the host was touched);  device_out() -> device copy becomes newest.
// Weight tensors additionally cache their repacked device form (gtb_weight_t), rebuilt when the host copy was touched.
#pragma once
#include <cstdint>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../gten_b200.h"
#include "gten_types.h"
#include "quants.h"

namespace gten {

inline int64_t& tensor_mem_allocated() { static int64_t v = 0; return v; }
#define G_TensorMemAllocated (::gten::tensor_mem_allocated())      // reference: a header-static counter (tensor.h:17)

class Tensor {
public:
    Tensor() = default;
    Tensor(const std::vector<int>& shape, Dtype dtype) : dtype_{dtype} {
        validate_shape(shape);
        shape_ = shape;
        set_strides_from_shape(shape);
        numel_ = numel_from_shape(shape);
        // byte sizing as the reference: quantised 2-D/3-D tensors are sized in whole blocks per row (tensor.cpp:37-58)
        const int ndim = (int)shape.size();
        if (dtype == kQint8 && ndim > 1) {
            const int rows = (ndim == 2) ? shape[0] : shape[0] * shape[1];
            storage_size_ = rows * ((shape[ndim - 1] + 31) / 32) * (int)sizeof(Q8Block);
        } else if (dtype == kQint4) {
            GTEN_ASSERT(ndim > 1 && shape[ndim - 1] % 32 == 0);
            const int rows = (ndim == 2) ? shape[0] : shape[0] * shape[1];
            storage_size_ = rows * (shape[ndim - 1] / 32) * (int)sizeof(Q4Block);
        } else {
            storage_size_ = numel_ * itemsize();
        }
        buf_ = std::make_shared<Buffer>((size_t)storage_size_, nullptr);
        G_TensorMemAllocated += storage_size_;
    }
    // non-owning view of caller memory (token ids, tinyllama.cpp:406): never freed, re-uploaded at every use
    Tensor(const void* data_ptr, const std::vector<int>& shape, Dtype dtype) : dtype_{dtype} {
        GTEN_ASSERTM(data_ptr != nullptr, "Expected a non-null pointer but got a nullptr.");
        validate_shape(shape);
        shape_ = shape;
        set_strides_from_shape(shape);
        numel_ = numel_from_shape(shape);
        storage_size_ = numel_ * itemsize();
        buf_ = std::make_shared<Buffer>((size_t)storage_size_, const_cast<void*>(data_ptr));
    }

    template <typename T> T* data_ptr() { return reinterpret_cast<T*>(host_rw()); }
    template <typename T> const T* data_ptr() const { return reinterpret_cast<const T*>(host_ro()); }
    void* data_ptr() { return host_rw(); }
    const void* data_ptr() const { return host_ro(); }

    Dtype dtype() const { return dtype_; }
    int itemsize() const {
        switch (dtype_) {
            case kQint8: return 1;
            case kInt32: return 4;
            case kFloat16: return 2;
            case kFloat32: return 4;
            default: GTEN_ASSERT(false); return 4;          // no Q4 case, like the reference (tensor.h:59-73)
        }
    }
    bool is_quantized() const { return dtype_ == kQint8; }
    bool is_1d() const { return shape_.size() == 1; }
    bool is_2d() const { return shape_.size() == 2; }
    bool is_3d() const { return shape_.size() == 3; }
    int ndims() const { return (int)shape_.size(); }
    int numel() const { return numel_; }
    int dimsize(int i) const { GTEN_ASSERT(i < (int)shape_.size()); return shape_[i]; }
    int stride(int i) const { GTEN_ASSERT(i < (int)strides_.size()); return strides_[i]; }
    int bstride(int i) const {
        GTEN_ASSERT(i < (int)strides_.size());
        if (dtype_ == kQint4) return strides_[i] == 1 ? 1 : (strides_[i] / 32) * (int)sizeof(Q4Block);
        if (dtype_ == kQint8) return strides_[i] == 1 ? 1 : (strides_[i] / 32) * (int)sizeof(Q8Block);
        return strides_[i] * itemsize();
    }
    size_t nbytes() const { return (size_t)storage_size_; }
    const std::vector<int>& shape() const { return shape_; }
    bool shape_eq(const std::vector<int>& s) const { return s == shape_; }

    // new shape over the same buffer; capacity is not checked beyond the element count (tensor.cpp:124-134)
    void resize(const std::vector<int>& new_shape) {
        validate_shape(new_shape);
        GTEN_ASSERT(new_shape.size() == shape_.size());
        shape_ = new_shape;
        set_strides_from_shape(new_shape);
        numel_ = numel_from_shape(new_shape);
    }
    Tensor view(const std::vector<int>& new_shape) const {
        Tensor t = *this;
        GTEN_ASSERT(numel_from_shape(new_shape) == numel_);
        t.shape_ = new_shape;
        t.set_strides_from_shape(new_shape);
        return t;
    }
    // reorders shape and strides of THIS tensor and returns it (the reference mutates too, tensor.cpp:173-190)
    Tensor permute(const std::vector<int>& order) {
        GTEN_ASSERT(order.size() == shape_.size());
        std::vector<int> s(shape_.size()), st(shape_.size());
        for (size_t i = 0; i < order.size(); i++) { s[i] = shape_[order[i]]; st[i] = strides_[order[i]]; }
        shape_ = s; strides_ = st;
        return *this;
    }
    void set_strides(const std::vector<int>& strides) { GTEN_ASSERT(strides.size() == shape_.size()); strides_ = strides; }
    std::string shape_str() const { return vec_str(shape_); }
    std::string strides_str() const { return vec_str(strides_); }
    void save(const std::string& path) const {
        std::ofstream f(path, std::ios::binary);
        GTEN_ASSERT(f.is_open());
        f.write(reinterpret_cast<const char*>(host_ro()), storage_size_);
    }
    void print_info() const {
        std::cout << "Tensor(shape=" << shape_str() << ", strides=" << strides_str() << ", dtype=" << dtype_str(dtype_)
                  << ", numel=" << numel_ << ", nbytes=" << storage_size_ << ", storage=HBM)\n";
    }
    void print() const {
        print_info();
        if (dtype_ == kFloat32 || dtype_ == kInt32) {
            const int n = numel_ < 16 ? numel_ : 16;
            for (int i = 0; i < n; i++) {
                if (dtype_ == kFloat32) std::cout << data_ptr<float>()[i] << ' ';
                else std::cout << data_ptr<Int32>()[i] << ' ';
            }
            std::cout << (numel_ > n ? "...\n" : "\n");
        }
    }
    friend std::ostream& operator<<(std::ostream& os, const Tensor& t) { t.print(); return os; }

    // ---- device side (used by gten::ops)
    const void* device_in() const { GTEN_ASSERT(buf_); buf_->to_device(); return buf_->dev; }
    void* device_out() { GTEN_ASSERT(buf_); buf_->to_device(); buf_->dev_newer = true; buf_->weight_stale = true; return buf_->dev; }
    gtb_weight_t weight_handle() const {
        GTEN_ASSERT(buf_ && is_2d());
        Buffer& b = *buf_;
        if (b.dev_newer) b.to_host();                       // a weight that was computed on the device (never happens on the path)
        if (b.weight == nullptr || b.weight_stale) {
            if (b.weight) gtb_weight_free(b.weight);
            b.weight = nullptr;
            GTEN_CUDA_OK(gtb_weight_upload(&b.weight, b.host_ptr(), (int)dtype_, shape_[0], shape_[1]));
            b.weight_stale = false;
        }
        return b.weight;
    }

private:
    struct Buffer {
        size_t nbytes;
        void* dev = nullptr;
        uint8_t* host = nullptr;          // owned mirror (lazily allocated)
        void* ext = nullptr;              // caller memory of a non-owning tensor
        bool host_touched = false;        // host copy may be newer than the device copy
        bool dev_newer = false;           // device copy is newer than the host copy
        bool weight_stale = true;
        gtb_weight_t weight = nullptr;
        Buffer(size_t n, void* external) : nbytes{n}, ext{external} {
            GTEN_CUDA_OK(gtb_malloc(&dev, n));
            if (!ext) GTEN_CUDA_OK(gtb_memset(dev, 0, n));
        }
        ~Buffer() {
            if (weight) gtb_weight_free(weight);
            if (dev) gtb_free(dev);
            delete[] host;
        }
        Buffer(const Buffer&) = delete;
        Buffer& operator=(const Buffer&) = delete;
        void* host_ptr() {
            if (ext) return ext;
            if (!host) host = new uint8_t[nbytes ? nbytes : 1]();      // a fresh buffer is zero on both sides
            return host;
        }
        void to_host() {
            void* h = host_ptr();
            if (dev_newer && !ext) { GTEN_CUDA_OK(gtb_d2h(h, dev, nbytes)); dev_newer = false; }
        }
        void to_device() {
            if (ext) { GTEN_CUDA_OK(gtb_h2d(dev, ext, nbytes)); GTEN_CUDA_OK(gtb_sync()); return; }   // caller memory may have changed
            if (host_touched && !dev_newer) { GTEN_CUDA_OK(gtb_h2d(dev, host, nbytes)); GTEN_CUDA_OK(gtb_sync()); }
            host_touched = false;
        }
    };

    void* host_rw() { GTEN_ASSERT(buf_); buf_->to_host(); buf_->host_touched = true; buf_->weight_stale = true; return buf_->host_ptr(); }
    const void* host_ro() const { GTEN_ASSERT(buf_); buf_->to_host(); return buf_->host_ptr(); }

    void validate_shape(const std::vector<int>& shape) const {
        GTEN_ASSERTM(shape.size() >= 1 && shape.size() <= 3, "Expected a 1-3 dimensional shape but got %d dims.", (int)shape.size());
        for (int d : shape) GTEN_ASSERTM(d > 0, "The value of dimension must be positive, got %d.", d);
    }
    void set_strides_from_shape(const std::vector<int>& shape) {
        strides_.assign(shape.size(), 1);
        for (int i = (int)shape.size() - 2; i >= 0; i--) strides_[i] = strides_[i + 1] * shape[i + 1];
    }
    int numel_from_shape(const std::vector<int>& shape) const { int n = 1; for (int d : shape) n *= d; return n; }
    static std::string vec_str(const std::vector<int>& v) {
        std::string s = "(";
        for (size_t i = 0; i < v.size(); i++) { s += std::to_string(v[i]); if (i + 1 < v.size()) s += ", "; }
        return s + ")";
    }

    Dtype dtype_ = kFloat32;
    std::shared_ptr<Buffer> buf_;
    int storage_size_ = 0;
    int numel_ = 0;
    std::vector<int> shape_;
    std::vector<int> strides_;
};

}  // namespace gten
