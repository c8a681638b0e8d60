// TEST INFRASTRUCTURE: <cuda_pipeline.h> for the emulated CTA.  Copies are DEFERRED like cp.async: a
// __pipeline_memcpy_async is only carried out by the __pipeline_wait_prior that retires its batch, so a
// kernel that reads a staged buffer without waiting for it reads stale data here as it would on the GPU.
#pragma once
#include <cstring>
#include <vector>
#include "../cta_emul.h"

namespace cta {
struct PendingCopy { void* dst; const void* src; size_t bytes; long batch; };
static std::vector<PendingCopy> g_pending[MAX_THREADS];
static long g_batch[MAX_THREADS];
}
static inline void __pipeline_memcpy_async(void* dst, const void* src, size_t bytes)
{
    cta::g_pending[cta::g.cur].push_back(cta::PendingCopy{dst, src, bytes, cta::g_batch[cta::g.cur]});
}
static inline void __pipeline_commit() { ++cta::g_batch[cta::g.cur]; }
static inline void __pipeline_wait_prior(int n)
{
    auto& q = cta::g_pending[cta::g.cur];
    const long done_below = cta::g_batch[cta::g.cur] - n;
    size_t keep = 0;
    for (size_t i = 0; i < q.size(); ++i) {
        if (q[i].batch < done_below) {
            if (q[i].src) memcpy(q[i].dst, q[i].src, q[i].bytes);
            else memset(q[i].dst, 0, q[i].bytes);          // cp.async with a source size of 0: zero fill
        } else q[keep++] = q[i];
    }
    q.resize(keep);
}

// the raw-PTX cp.async of csrc/backx_kernels.cuh (16 bytes, zero fill when the source size is 0)
#define AXS_CP_ASYNC16(dst, src, bytes) \
    cta::g_pending[cta::g.cur].push_back(cta::PendingCopy{(void*)(dst), (bytes) ? (const void*)(src) : nullptr, 16, cta::g_batch[cta::g.cur]})
#define AXS_CP_ASYNC_COMMIT() __pipeline_commit()
#define AXS_CP_ASYNC_WAIT(n) __pipeline_wait_prior(n)
