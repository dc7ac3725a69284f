// plan_cost.cpp -- what a plan-cache MISS costs on the batcher thread (host only, no GPU): the streaming geometry of the
// resize + watermark pass and of the crop thumbnail for the source sizes of BASELINE's mixed stream, built cold.
//   g++ -O2 -std=c++17 -ffp-contract=off -o plan_cost plan_cost.cpp ../../imageprocessor_b200/csrc/plan.cpp && ./plan_cost
#include <algorithm>
#include <chrono>
#include <cstdio>

#include "../../imageprocessor_b200/csrc/plan.h"

using namespace ipg;
using clk = std::chrono::steady_clock;

static double ms(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }

int main()
{
    const int sizes[][2] = {{640, 480}, {1632, 1224}, {2832, 2124}, {4000, 3000}, {6656, 4992}, {7680, 4320}, {8000, 6000}};
    printf("[\n");
    bool first = true;
    for (auto &wh : sizes) {
        const int W = wh[0], H = wh[1];
        // the reference's geometry: keep-aspect into 1024 x 768, centre crop square to 200
        const double ratio = std::min(1024.0 / W, 768.0 / H);
        const int nw = (int)(W * ratio), nh = (int)(H * ratio), cs = std::min(W, H);
        const StreamTargetSpec rs{0, 0, W, H, nw, nh}, th{(W - cs) / 2, (H - cs) / 2, cs, cs, 200, 200};
        double t_rs = 1e9, t_th = 1e9;
        int vint = 0;
        size_t recs = 0;
        for (int rep = 0; rep < 5; rep++) { // a different band hint each time: a different cache key, the axis tables stay cached
            const int bands = 3 + rep;
            auto t0 = clk::now();
            auto g1 = get_stream_geom(W, H, &rs, 1, true, bands, 257.0);
            auto t1 = clk::now();
            auto g2 = get_stream_geom(W, H, &th, 1, false, bands, 257.0);
            auto t2 = clk::now();
            if (g1) t_rs = std::min(t_rs, ms(t0, t1));
            if (g2) { t_th = std::min(t_th, ms(t1, t2)); vint = g2->vint_ok; recs = g2->grec.size(); }
        }
        printf("%s {\"source\": \"%dx%d\", \"resize_watermark_plan_ms\": %.3f, \"thumbnail_plan_ms\": %.3f, \"thumbnail_integer_form\": %s, \"thumbnail_group_records\": %zu}",
               first ? "" : ",\n", W, H, t_rs > 1e8 ? -1.0 : t_rs, t_th > 1e8 ? -1.0 : t_th, vint ? "true" : "false", recs);
        first = false;
    }
    printf("\n]\n");
    return 0;
}
