// rbc_common.h — error plumbing shared by the translation units of librbc_b200.so
#pragma once
#include <cuda_runtime.h>

#include <string>

int rbc_fail(const std::string& m);   // records the message for rbc_last_error() and returns -1

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return rbc_fail(std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    } while (0)

// ------------------------------------------------------------------------------------------
// Host-buffer step entry points (rbc2d_step_host / rbc3d_step_host): the batch is launched in up to kMaxChunks chunks
// of whole waves of the persistent grid, and a second stream copies each chunk's outputs to the caller's buffers while
// the next chunk computes — the device-to-host transfer of a step (19 MB in 2D, 310 MB in 3D at the bench sizes)
// overlaps the kernel instead of following it.
// ------------------------------------------------------------------------------------------
struct HostPipe {
    static constexpr int kMaxChunks = 4;
    cudaStream_t copy = nullptr;
    cudaEvent_t done[kMaxChunks] = {};
    int* iota = nullptr;             // device: 0, 1, ..., B-1 (env_ids of a chunk = iota + offset)
    int n = 0;
};
int rbc_pipe_prepare(HostPipe* p, int B);          // lazily creates the stream, the events and iota
void rbc_pipe_destroy(HostPipe* p);
// envs of chunk c when B environments run on `lanes` concurrent CTAs / clusters; returns the number of chunks
inline int rbc_pipe_chunks(int B, int lanes, int* len /*[kMaxChunks]*/)
{
    const int waves = (B + lanes - 1) / lanes;
    const int nch = waves < HostPipe::kMaxChunks ? (waves < 1 ? 1 : waves) : HostPipe::kMaxChunks;
    int off = 0;
    for (int c = 0; c < nch; ++c) {
        const int w = waves / nch + (c < waves % nch ? 1 : 0);
        int l = w * lanes;
        if (l > B - off) l = B - off;
        len[c] = l;
        off += l;
    }
    return nch;
}

// ------------------------------------------------------------------------------------------
// episode bookkeeping of the fused vector environments (rbc2d_vec_* / rbc3d_vec_*), implemented in rbc2d_lib.cu
// ------------------------------------------------------------------------------------------
// episode += 1 (restart: := 1), episode return := 0, pending := 0 for the listed environments (env_ids NULL = the first n)
int rbc_vec_mark(cudaStream_t stream, const int* env_ids, int n, long long* episode, double* ep_return, int* pending, int restart);
// idx[env] = checkpoint draw of episode 0 for env = 0..n-1
int rbc_vec_draw(cudaStream_t stream, int* idx, int n, unsigned long long seed, unsigned long long offset, int n_ep);
