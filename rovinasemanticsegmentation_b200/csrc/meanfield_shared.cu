// The point kernel of the fused mean-field iteration, built for an SM it shares with a CTA of the cooperative blur: at most
// RSS_POINT_SHARED_NREG = 64 registers (see meanfield_point.inl).  Used while several keyframes are in flight on the GPU.
#include "kernels.hpp"
#include "lattice.cuh"
#include "meanfield.cuh"

#ifndef RSS_POINT_SHARED_NREG
#define RSS_POINT_SHARED_NREG 64
#endif
#define RSS_POINT_BOUNDS __maxnreg__(RSS_POINT_SHARED_NREG)
#define RSS_POINT_NS point_shared
#define RSS_POINT_ENTRY launch_meanfield_fused_shared
#include "meanfield_point.inl"

namespace rss {
cudaError_t launch_meanfield_fused_shared(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary,
                                          float* Q, uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode) {
    return point_shared::launch_meanfield_fused_shared(c, st, a, d1a, d1b, unary, Q, labels, tm, G, ls, mode);
}
}  // namespace rss
