// adapter/Exchange.cc -- see Exchange.h.  Plain host code: the kernel lives in libgsb.so.
#include "Exchange.h"

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/distributed/c10d/symm_mem/SymmetricMemory.hpp>

#include "gsb.h"

namespace ORB_SLAM2 {

namespace symm = c10d::symmetric_memory;

GradientExchange::GradientExchange(int64_t capacity_floats, c10::Device device, const std::string& group_name)
{
    TORCH_CHECK(device.is_cuda(), "GradientExchange needs a CUDA device");
    capacity_ = (capacity_floats + 3) / 4 * 4;
    c10::cuda::CUDAGuard guard(device);
    // the scratch size depends on the world size, which only the rendezvous tells: probe it with the largest world the kernel takes
    const int64_t sync_words = (int64_t)gsb_exchange_sync_bytes(16) / 4;
    const int64_t total = capacity_ + sync_words;
    buf_ = symm::empty_strided_p2p({total}, {1}, c10::ScalarType::Float, device, group_name, std::nullopt);
    buf_.zero_();
    auto hdl = symm::rendezvous(buf_, group_name);
    rank_ = hdl->get_rank();
    world_ = hdl->get_world_size();
    TORCH_CHECK(world_ >= 2 && world_ <= 16, "GradientExchange: world size ", world_, " outside [2, 16]");
    for (void* p : hdl->get_buffer_ptrs()) {
        peers_.push_back(p);
        sync_.push_back(static_cast<float*>(p) + capacity_);
    }
    multicast_ = hdl->has_multicast_support() ? hdl->get_multicast_ptr() : nullptr;
    c10::cuda::getCurrentCUDAStream(device.index()).synchronize();
    hdl->barrier(0, 60000);   // every rank's scratch is zeroed before anyone handshakes
    handle_ = hdl;
}

torch::Tensor GradientExchange::alloc(int64_t n)
{
    const int64_t n4 = (n + 3) / 4 * 4;
    TORCH_CHECK(n >= 0 && used_ + n4 <= capacity_, "GradientExchange: symmetric allocation exhausted (", used_, " + ", n4, " > ", capacity_, ")");
    torch::Tensor t = buf_.narrow(0, used_, n);
    used_ += n4;
    return t;
}

void GradientExchange::allreduce(const torch::Tensor& t, bool use_multicast)
{
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == c10::ScalarType::Float && t.is_contiguous(), "GradientExchange: contiguous CUDA fp32 tensor expected");
    const int64_t off = (static_cast<const char*>(t.data_ptr()) - static_cast<const char*>(buf_.data_ptr()));   // bytes
    const int64_t n4 = (t.numel() + 3) / 4 * 4;   // 16-byte words: an odd-P block is rounded up into the zero padding behind it
    TORCH_CHECK(off >= 0 && off % 16 == 0 && off + n4 * 4 <= capacity_ * 4, "GradientExchange: tensor is not a 16-byte-aligned slice of the symmetric allocation");
    std::vector<void*> peer(world_);
    for (int r = 0; r < world_; r++) peer[r] = static_cast<char*>(peers_[r]) + off;
    void* mc = (multicast_ && use_multicast) ? static_cast<char*>(multicast_) + off : nullptr;
    c10::cuda::CUDAGuard guard(t.device());
    const int rc = gsb_exchange_allreduce(mc, peer.data(), sync_.data(), n4, rank_, world_, (gsb_stream_t)c10::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == GSB_OK, "gsb_exchange_allreduce: ", gsb_last_error());
}

}  // namespace ORB_SLAM2
