"""cmax_slam_b200 -- B200-native (sm_100a) contrast-maximisation inner loop of CMax-SLAM.

The product is libcmax_b200.so (hand-written CUDA + a plain C ABI, include/cmax_b200.h).  This
package only holds what that path needs: the CUDA / C++ sources (csrc/), the in-tree build, the ctypes
binding (_capi), the host-side mirrors of the reference's interfaces -- cost functions (frontend / backend),
trajectory initialisation (trajectory), event ingestion and the bearing-vector LUT (stream), the node's wiring
(pipeline) -- the synthetic workloads of BASELINE.json (synth) and the multi-GPU sharding drivers (dist).
There is no CPU fallback: without the library or a CUDA device the device-side constructors raise.
"""
__all__ = ["frontend", "backend", "trajectory", "stream", "pipeline", "synth", "dist", "build"]
