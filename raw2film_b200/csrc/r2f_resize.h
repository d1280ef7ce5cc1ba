// Device cv2.resize (INTER_AREA / INTER_LANCZOS4) as the reference's resolution_scaling uses it; see r2f_resize.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

namespace r2f {

struct AreaTabHost {
    std::vector<int> start;    // dsize + 1 offsets into si / alpha
    std::vector<int> si;       // source index per tap
    std::vector<float> alpha;  // float32 weight per tap
};
struct AreaTabDev {
    const int *start;
    const int *si;
    const float *alpha;
};
struct LanczosTabHost {
    std::vector<int> ofs;      // floor of the source coordinate per destination index
    std::vector<float> coef;   // dsize x 8 float32 weights
    std::vector<int> icoef;    // the same in 1/2048 fixed point (uint8 images)
};
struct LanczosTabDev {
    const int *ofs;
    const float *coef;
    const int *icoef;
};

bool resize_area_is_fast(int ssize, int dsize, int &iscale);
void resize_area_tab(int ssize, int dsize, AreaTabHost &t);
void resize_lanczos4_tab(int ssize, int dsize, LanczosTabHost &t);

cudaError_t launch_resize_area(const void *src, bool u8, int cin, int sh, int sw, void *dst, int dh, int dw,
                               const AreaTabDev &xt, const AreaTabDev &yt, int num_sms, cudaStream_t st);
cudaError_t launch_resize_area_int(const void *src, bool u8, int cin, int sh, int sw, void *dst, int dh, int dw, int ix,
                                   int iy, int num_sms, cudaStream_t st);
cudaError_t launch_resize_lanczos4(const void *src, bool u8, int cin, int sh, int sw, void *dst, int dh, int dw,
                                   const LanczosTabDev &xt, const LanczosTabDev &yt, int num_sms, cudaStream_t st);

}  // namespace r2f
