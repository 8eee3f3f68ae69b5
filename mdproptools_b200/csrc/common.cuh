// Shared host-side plumbing for libmdprop_b200: context, scratch arena, error reporting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mdprop_b200.h"

#define MDP_ERR_CUDA (-1)
#define MDP_ERR_ARG (-2)
#define MDP_ERR_OOM (-3)
#define MDP_ERR_IO (-4)

void mdp_set_error(const char *fmt, ...);

#define MDP_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            mdp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return MDP_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define MDP_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            mdp_set_error(__VA_ARGS__);                                                             \
            return MDP_ERR_ARG;                                                                     \
        }                                                                                           \
    } while (0)

// A bump arena over one cudaMalloc'ed slab.  Calls carve their scratch from it at entry; because all
// work of one context is stream ordered on the caller's stream, the slab can be reused by the next
// call without synchronising.  Growing the slab synchronises the device once.
struct mdp_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    char *slab = nullptr;
    size_t slab_bytes = 0;
    size_t slab_used = 0;
    size_t slab_limit = (size_t)8 << 30;
    int64_t launches = 0;
    // last pair call statistics (device side, 4 x int64) and pinned host mirror
    unsigned long long *d_stats = nullptr;

    // optional per-kernel timing (bench.py): CUDA event pairs recorded on the launching stream
    struct Timed {
        int tag;
        cudaEvent_t e0, e1;
    };
    bool timing = false;
    std::vector<Timed> timed;
    cudaEvent_t timer_begin(int tag, cudaStream_t st)
    {
        if (!timing) return nullptr;
        Timed t;
        t.tag = tag;
        cudaEventCreate(&t.e0);
        cudaEventCreate(&t.e1);
        cudaEventRecord(t.e0, st);
        timed.push_back(t);
        return t.e1;
    }
    void timer_end(cudaEvent_t e1, cudaStream_t st)
    {
        if (e1) cudaEventRecord(e1, st);
    }

    void arena_reset() { slab_used = 0; }
    int arena_reserve(size_t bytes);              // make sure the slab holds at least `bytes`
    void *arena_take(size_t bytes)
    {
        size_t off = (slab_used + 255) & ~(size_t)255;
        if (off + bytes > slab_bytes) return nullptr;
        slab_used = off + bytes;
        return slab + off;
    }
};

static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

template <class T>
static inline T ceil_div(T a, T b)
{
    return (a + b - 1) / b;
}

#define MDP_LAUNCHED(ctx) ((ctx)->launches++)

static inline int mdp_check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        mdp_set_error("kernel launch %s failed: %s", what, cudaGetErrorString(e));
        return MDP_ERR_CUDA;
    }
    return 0;
}
