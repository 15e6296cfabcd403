// CGD per-group correlation (Gram) loss - builder-defined extension.
// Placeholder entry points; the tcgen05/TMEM kernel lands in a later milestone.
#include <cuda_runtime.h>

#include "../../include/segdistill.h"

extern "C" {

size_t sd_cgd_corr_workspace_bytes(int B, int C, int HW, int group) {
    (void)B; (void)C; (void)HW; (void)group;
    return 256;
}

int sd_cgd_corr_fwd_bwd(const void* S, const void* T, void* dS, float* loss, int B, int C, int HW, int group,
                        int dtype, float alpha, float grad_scale, void* workspace, size_t workspace_bytes,
                        void* stream) {
    (void)S; (void)T; (void)dS; (void)loss; (void)B; (void)C; (void)HW; (void)group; (void)dtype;
    (void)alpha; (void)grad_scale; (void)workspace; (void)workspace_bytes; (void)stream;
    return SD_ERR_UNSUPPORTED;
}

}  // extern "C"
