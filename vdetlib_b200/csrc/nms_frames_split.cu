// nms_frames_split.cu -- the two-array sort variants of nms_frames_kernel (nms_frames.cuh): frames that fill at
// most 5/8 of a power-of-two sort network (129..160 and 257..320 boxes; BASELINE's 300-box frames sort
// 256 + 64 keys).  A translation unit of its own so that the instantiations compile beside the others.
#include "nms_frames.cuh"

namespace vdet {

template <int NPER, int NPB>
static int launch_split(int threads, const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    if (threads == NMS_THREADS_WIDE) return launch_nms_frames_t<NPER, NPB, true, NMS_THREADS_WIDE>(p, smem, grid, st);   // staged only
    return p.stage ? launch_nms_frames_t<NPER, NPB, true>(p, smem, grid, st)
                   : launch_nms_frames_t<NPER, NPB, false>(p, smem, grid, st);
}

int launch_nms_frames_split(int nper, int npb, int threads, const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    if (nper == 4 && npb == 1) return launch_split<4, 1>(threads, p, smem, grid, st);
    if (nper == 8 && npb == 2) return launch_split<8, 2>(threads, p, smem, grid, st);
    set_error("nms_frames: no two-array variant for %d + %d keys per lane", nper, npb);
    return VDET_ERR_UNSUPPORTED;
}

}  // namespace vdet
