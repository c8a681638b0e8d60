// TEST INFRASTRUCTURE: <cooperative_groups.h> for the emulated CTA - a cluster of ONE thread block (the
// CLUSTER = false instances of the kernels; distributed shared memory is not emulated).
#pragma once
namespace cooperative_groups {
struct cluster_group {
    unsigned num_blocks() const { return 1; }
    unsigned block_rank() const { return 0; }
    void sync() const {}
    template <class T> T* map_shared_rank(T* p, unsigned) const { return p; }
};
static inline cluster_group this_cluster() { return cluster_group(); }
}
