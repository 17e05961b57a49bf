// nj_loop.h -- internal interface between the host driver of the NJ phase (nj_host.cpp) and the device-resident join
// loop (nj_loop_logic.h).  Two implementations with the same entry points: the CUDA one inside libvft_b200.so
// (vft_cuda.cu / nj_loop_gpu.cuh: state in HBM, one thread block runs the logic, the grid evaluates the distances)
// and the one-thread CPU double inside oracle/libvftoracle.so (oracle/nj_loop_cpu.cpp: the same logic source over the
// CPU restatement of the kernels), which exists so that the logic can be tested without a GPU.
//
// Life cycle: create -> upload(image after setAllLeafTopHits + the first resetTopVisible) -> run ... -> download.
// run() returns when the loop is DONE (three active nodes left) or when it needs an operation this backend leaves to the
// caller (status NEED_RESET / NEED_REFRESH: the O(nActive) rebuild of the top-visible set, a top-hits refresh); the caller
// downloads the image, performs it with the host code, uploads and runs again.
#pragma once
#include <cstdint>
#include "../../include/vft_b200.h"

extern "C" {

typedef struct vftx_loop vftx_loop;

typedef struct vftx_loop_image {             // host memory, caller-owned; numeric_t arrays as void*
    int64_t nSeqs, maxnodes, m, nTV;
    int64_t maxnode, nActive, topvisibleAge, nActiveOutProfileReset;
    double totdiam;
    int32_t *parent, *up, *child;            // [M], [M], [M*3]
    void *branchlength, *diameter, *outDist; // [M]
    int32_t *nOutAct;                        // [M]
    int32_t *hitJ; void *hitDist;            // [M*m]
    int32_t *hitCount, *age;                 // [M]
    int32_t *visJ; void *visDist;            // [M]
    int32_t *topvisible;                     // [nTV]
} vftx_loop_image;

typedef struct vftx_loop_status {
    int32_t status;                          // njl::ST_*
    int32_t resume;                          // njl::RS_*
    int32_t visfixPending;                   // NEED_RESET: the visible-set repair loop of NJ.tcc:4171-4201 comes first
    int32_t newnode;                         // NEED_REFRESH: the node whose topHitJoin asked for it
    int64_t nActive, maxnode;
    int64_t nJoins, nRefresh, nVisibleUpdate, nHillBetter, nReset, nInlineOut, nInlinePair, nPairHit, nRebuild, nSteps;
} vftx_loop_status;

int vftx_loop_create(vft_ctx *ctx, const vft_nj_options *opt, int64_t m, int64_t nTV, vftx_loop **out);
int vftx_loop_upload(vftx_loop *lp, const vftx_loop_image *img, int32_t resume);
int vftx_loop_run(vftx_loop *lp, vftx_loop_status *st);
int vftx_loop_download(vftx_loop *lp, vftx_loop_image *img);
// joins made by the loop so far, (i, j) pairs in order (for vft_nj_result.joins); returns the count
int64_t vftx_loop_joins(vftx_loop *lp, int64_t *out, int64_t maxJoins);
int vftx_loop_destroy(vftx_loop *lp);

}
