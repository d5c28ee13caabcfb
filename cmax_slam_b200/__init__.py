"""cmax_slam_b200 -- B200-native (sm_100a) contrast-maximisation inner loop of CMax-SLAM.

The product is libcmax_b200.so (hand-written CUDA + a plain C ABI, include/cmax_b200.h).  This
package only holds what that path needs: the CUDA sources (csrc/), the in-tree build, the ctypes
binding, the host-side mirrors of the reference's cost-function interface (frontend / backend),
the synthetic workloads of BASELINE.json (synth) and the hypothesis-sharding driver (dist).
There is no CPU fallback: without the library or a CUDA device the constructors raise.
"""
__all__ = ["frontend", "backend", "synth", "dist", "build"]
