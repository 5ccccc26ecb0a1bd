// adapter/Exchange.h -- C++ (libtorch) host of the ONE exchange step of the sharded render-optimise loop: the sum all-reduce of
// the packed [14, P] gradient block between loss.backward() and the Adam step (src/Render.cc:471-475), run as ONE libgsb
// kernel over NVLink / NVSwitch peer memory (gsb_exchange_allreduce, csrc/exchange.cu) instead of an NCCL call.
//
// The reference is single-GPU, so this class has no counterpart there: it is what a maintainer adds next to Gaussian.cc when
// the keyframe-batch shard of INTEGRATION.md is enabled (one process per GPU, a c10d process group registered under
// `group_name`).  Twin of gsorb_slam_b200/distributed.py::SymmetricExchange; both only hand device pointers to the C ABI.
//
//   GradientExchange xch(14 * P, device, group_name);     // collective: symmetric allocation + rendezvous
//   torch::Tensor block = xch.alloc(14 * P);               // the [14, P] block lives INSIDE the symmetric allocation
//   ... gsb_backward / gsb_prologue_backward write their rows into views of `block` ...
//   xch.allreduce(block);                                   // in place, on the current CUDA stream
#pragma once
#include <torch/torch.h>

#include <string>
#include <vector>

namespace ORB_SLAM2 {

class GradientExchange {
public:
    // Collective over the ranks of `group_name`: allocates capacity_floats (+ the handshake scratch) of symmetric, peer-mapped
    // memory on `device`, zeroes it and exchanges the mappings.  Throws c10::Error when symmetric memory is unavailable.
    GradientExchange(int64_t capacity_floats, c10::Device device, const std::string& group_name);

    // fp32 tensor of n elements inside the symmetric allocation (same offset on every rank when all ranks call in the same
    // order); reservations are rounded up to 16 bytes, the padding stays zero.
    torch::Tensor alloc(int64_t n);

    // In-place sum over the ranks of a tensor obtained from alloc(), on the current CUDA stream.  use_multicast: reduce through the
    // NVSwitch multicast mapping (multimem.ld_reduce / multimem.st) when the box offers one, else 128-bit peer loads / stores.
    void allreduce(const torch::Tensor& t, bool use_multicast = true);

    int rank() const { return rank_; }
    int world_size() const { return world_; }
    bool has_multicast() const { return multicast_ != nullptr; }

private:
    torch::Tensor buf_;                 // [data | handshake scratch]
    std::vector<void*> peers_, sync_;   // every rank's mapping of the data / of the scratch
    void* multicast_ = nullptr;
    int64_t capacity_ = 0, used_ = 0;
    int rank_ = 0, world_ = 1;
    c10::intrusive_ptr<c10::intrusive_ptr_target> handle_;   // keeps the c10d::symmetric_memory mapping alive
};

}  // namespace ORB_SLAM2
