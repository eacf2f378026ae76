/* A plain C program against include/fdms.h on the GPU box (tests/test_gpu_api.py compiles and runs it):
 * fdms_create -> device buffers through the CUDA runtime -> fdms_nms -> results back -> fdms_destroy.
 * No Python, no torch: what a non-Python maintainer of the reference would link.  Prints the keep list. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "fdms.h"

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main(void) {
    fdms_ctx *ctx = NULL;
    int rc = fdms_create(&ctx, 0);
    if (rc != FDMS_OK) { printf("fdms_create failed %d\n", rc); return 1; }
    /* 6 boxes, rows of 5 floats [x1 y1 x2 y2 score]: two clusters of heavily overlapping boxes and one loner */
    const float rows[6][5] = {{0, 0, 10, 10, 0.9f}, {1, 1, 11, 11, 0.8f}, {0, 0, 10, 10, 0.85f},
                              {50, 50, 60, 60, 0.7f}, {51, 50, 61, 60, 0.95f}, {100, 100, 110, 110, 0.3f}};
    const int32_t count = 6;
    float *d_rows; int32_t *d_count, *d_keep, *d_nkeep;
    CHECK(cudaMalloc((void **)&d_rows, sizeof(rows)));
    CHECK(cudaMalloc((void **)&d_count, 4));
    CHECK(cudaMalloc((void **)&d_keep, 6 * 4));
    CHECK(cudaMalloc((void **)&d_nkeep, 4));
    CHECK(cudaMemcpy(d_rows, rows, sizeof(rows), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_count, &count, 4, cudaMemcpyHostToDevice));
    cudaStream_t st;
    CHECK(cudaStreamCreate(&st));
    rc = fdms_nms(ctx, d_rows, 5, 0, 4, NULL, d_count, 1, 6, 0.5, FDMS_IOU_TV_CUDA, 0, d_keep, d_nkeep, (fdms_stream)st);
    if (rc != FDMS_OK) { printf("fdms_nms failed %d: %s\n", rc, fdms_last_error(ctx)); return 1; }
    CHECK(cudaStreamSynchronize(st));
    int32_t keep[6], nk = -1;
    CHECK(cudaMemcpy(&nk, d_nkeep, 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(keep, d_keep, 6 * 4, cudaMemcpyDeviceToHost));
    printf("n_keep %d keep", nk);
    for (int i = 0; i < nk; i++) printf(" %d", keep[i]);
    printf("\n");
    /* the host-side geometry entry point too (utils/datasets.py:873) */
    fdms_geom g;
    rc = fdms_letterbox_geometry(1080, 1920, 640, 640, 1, 0, 1, 32, &g);
    printf("letterbox %d %d %d top %d\n", rc, g.dst_h, g.dst_w, g.top);
    /* bad arguments are reported, not crashed on */
    rc = fdms_nms(ctx, NULL, 5, 0, 4, NULL, d_count, 1, 6, 0.5, FDMS_IOU_TV_CUDA, 0, d_keep, d_nkeep, (fdms_stream)st);
    printf("null rows -> %d (%s)\n", rc, fdms_last_error(ctx));
    cudaFree(d_rows); cudaFree(d_count); cudaFree(d_keep); cudaFree(d_nkeep);
    cudaStreamDestroy(st);
    fdms_destroy(ctx);
    return 0;
}
