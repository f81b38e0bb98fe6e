// Replay-batch assembly on the device.
//
// The reference keeps every replay frame as a CPU float tensor [3, H, W] in [0, 1]
// (datasets/img_utils.py:27-37 `img_np_toTensor`: uint8 HWC -> permute -> float / 255), stacks a batch of
// them in a Python loop (datasets/env_img_replay_buffer.py:68-116 `sample_random_batch_seq`) and copies
// 2 x B x 3 x H x W floats to the GPU every optimisation step (libero/lb_online_trainer_v7.py:586).
// Here the episodes live in HBM as the uint8 HWC frames the simulator rendered (49 KB per 128 x 128 frame,
// a quarter of the float copy; 180 GB of HBM holds every frame the trainer ever keeps), and a step's batch is
// one gather: per sample the device address of its frame -> fp32 [n, 3, H, W] = float(u8) / 255, bit for bit
// what the reference's tensors hold.  HBM-bound: 3 B read + 12 B written per pixel.
#include "common.cuh"
#include "../../include/v2a_b200.h"

#include <atomic>

namespace v2a {
extern std::atomic<int64_t> g_launches;

#define V2A_REPLAY_LAUNCH_OK()               \
    do {                                     \
        V2A_CUDA_OK(cudaGetLastError());     \
        v2a::g_launches.fetch_add(1);        \
    } while (0)

constexpr int kReplayThreads = 256;

// grid (ceil(H / rows), n).  One block = `rows` image rows of one sample: the rows' bytes are contiguous in the
// HWC frame, staged through shared memory with 16-byte loads, and leave as three contiguous fp32 runs (one per
// channel plane) with 16-byte stores.
__global__ void __launch_bounds__(kReplayThreads) replay_gather_images_kernel(
    const uint8_t* const* __restrict__ frames, int H, int W, int rows, float* __restrict__ out) {
    extern __shared__ uint4 stage4[];
    uint8_t* stage = reinterpret_cast<uint8_t*>(stage4);
    const int sample = blockIdx.y;
    const int h0 = blockIdx.x * rows;
    const int nrows = min(rows, H - h0);
    const int npix = nrows * W;
    const int nbytes = npix * 3;
    const uint8_t* src = frames[sample] + (size_t)h0 * W * 3;

    int done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n16 = nbytes >> 4;
        const uint4* src4 = reinterpret_cast<const uint4*>(src);
        for (int i = threadIdx.x; i < n16; i += kReplayThreads) stage4[i] = __ldg(src4 + i);
        done = n16 << 4;
    }
    for (int i = done + threadIdx.x; i < nbytes; i += kReplayThreads) stage[i] = __ldg(src + i);
    __syncthreads();

    const size_t plane = (size_t)H * W;
    float* dst = out + (size_t)sample * 3 * plane + (size_t)h0 * W;
    const bool vec = ((npix & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((plane & 3) == 0);
    if (vec) {
        const int nq = npix >> 2;
        for (int i = threadIdx.x; i < 3 * nq; i += kReplayThreads) {
            const int c = i / nq, q = i - c * nq;
            const uint8_t* s = stage + q * 12 + c;
            float4 v;
            v.x = __fdiv_rn((float)s[0], 255.0f);
            v.y = __fdiv_rn((float)s[3], 255.0f);
            v.z = __fdiv_rn((float)s[6], 255.0f);
            v.w = __fdiv_rn((float)s[9], 255.0f);
            reinterpret_cast<float4*>(dst + (size_t)c * plane)[q] = v;
        }
    } else {
        for (int i = threadIdx.x; i < 3 * npix; i += kReplayThreads) {
            const int c = i / npix, p = i - c * npix;
            dst[(size_t)c * plane + p] = __fdiv_rn((float)stage[p * 3 + c], 255.0f);
        }
    }
}

// out[b][t][a] = rows[b][t * A + a]: `rows[b]` = device address of the first of T consecutive action rows
__global__ void replay_gather_actions_kernel(const float* const* __restrict__ rows, int B, int TA,
                                             float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * TA) return;
    const int b = i / TA;
    out[i] = __ldg(rows[b] + (i - b * TA));
}

}  // namespace v2a

extern "C" {

int v2a_replay_gather_images(const void* const* frame_ptrs, int n, int H, int W, float* out, void* stream) {
    V2A_REQUIRE(frame_ptrs && out, "replay_gather_images: missing pointers");
    V2A_REQUIRE(n >= 0 && H > 0 && W > 0 && n <= 65535, "replay_gather_images: bad shape n=%d H=%d W=%d", n, H, W);
    if (n == 0) return 0;
    // about 3 KB of frame bytes per block, whole rows, at most 48 KB of shared memory
    int rows = (3072 + W * 3 - 1) / (W * 3);
    if (rows > H) rows = H;
    const size_t smem = (((size_t)rows * W * 3 + 15) / 16) * 16;
    V2A_REQUIRE(smem <= 48 * 1024, "replay_gather_images: image rows of %d pixels do not fit the staging buffer", W);
    dim3 grid((unsigned)v2a::ceil_div(H, rows), (unsigned)n);
    v2a::replay_gather_images_kernel<<<grid, v2a::kReplayThreads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint8_t* const*>(frame_ptrs), H, W, rows, out);
    V2A_REPLAY_LAUNCH_OK();
    return 0;
}

int v2a_replay_gather_actions(const void* const* row_ptrs, int B, int T, int A, float* out, void* stream) {
    V2A_REQUIRE(row_ptrs && out, "replay_gather_actions: missing pointers");
    V2A_REQUIRE(B >= 0 && T > 0 && A > 0, "replay_gather_actions: bad shape B=%d T=%d A=%d", B, T, A);
    if (B == 0) return 0;
    const int total = B * T * A;
    v2a::replay_gather_actions_kernel<<<(unsigned)v2a::ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float* const*>(row_ptrs), B, T * A, out);
    V2A_REPLAY_LAUNCH_OK();
    return 0;
}

}  // extern "C"
