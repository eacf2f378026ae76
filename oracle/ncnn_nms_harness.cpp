// Harness around the REFERENCE's own ncnn-demo NMS (cpp/yolov7-face-ncnn/src/yolov7face.cpp:35-74) so that
// tests/golden/make_golden.py can pin the oracle's mode 2 against it.  The two .inc files are cut out of the
// reference sources where they lie by oracle/Makefile (target _ref/ncnn_nms.so) into the git-ignored oracle/_ref/;
// nothing of the reference is copied into the repository.  Build container only.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "ncnn_types.inc"  // NUM_KEYPOINTS, struct Point, struct Object   (include/yolov7face.h)
#include "ncnn_nms.inc"    // bool cmp(Object, Object); void nms(std::vector<Object>&, float)

extern "C" int64_t ref_ncnn_nms(const float *boxes, const float *scores, int64_t n, float thresh, int64_t *keep) {
    std::vector<Object> v((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        v[i].x1 = boxes[4 * i];
        v[i].y1 = boxes[4 * i + 1];
        v[i].x2 = boxes[4 * i + 2];
        v[i].y2 = boxes[4 * i + 3];
        v[i].score = scores[i];
        v[i].landmark[0].x = (float)i;  // carries the identity through sort + nms (exact for n < 2^24)
    }
    std::sort(v.begin(), v.end(), cmp);  // yolov7face.cpp:191 (callers of this harness pass distinct scores)
    nms(v, thresh);                      // yolov7face.cpp:192
    for (size_t k = 0; k < v.size(); k++) keep[k] = (int64_t)v[k].landmark[0].x;
    return (int64_t)v.size();
}
