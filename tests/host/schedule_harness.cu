// Host-side check of the tile schedule the tcgen05 conv kernels walk (CPU test, no GPU): choose_schedule (y2_internal.h) +
// SegIter / CapIter (y2_ptx.cuh), compiled for the host.  For every configuration: every (tile, k-block) is visited exactly
// once over all workers; sub-segments are contiguous, no longer than the cap, equal-sized within a segment; a worker's
// non-head segment (the partial a stream-K head collects) is the FIRST segment of its stream-K phase (so it is published
// before any head can need it: no circular wait); a head with contributors is a worker's last segment (the shared-memory
// ring it stages the partials through is idle).
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../yolo_tf_b200/csrc/y2_internal.h"
#include "../../yolo_tf_b200/csrc/y2_ptx.cuh"

namespace y2 {
int g_sched_override = 0;
double g_sched_handoff_kb = 13.0;
}  // namespace y2

static int check(long long tiles, int KB, int workers, int max_ctas, double kbw, size_t shared_bytes, int cap, bool verbose) {
    int dp_tiles = 0, sk = 0, grid = 0;
    y2::choose_schedule(tiles, KB, workers, max_ctas, kbw, shared_bytes, &dp_tiles, &sk, &grid);
    if (grid < 1 || grid > workers || dp_tiles < 0 || dp_tiles > tiles || sk < 0 || sk > grid) {
        printf("FAIL schedule tiles=%lld KB=%d workers=%d: dp=%d sk=%d grid=%d\n", tiles, KB, workers, dp_tiles, sk, grid);
        return 1;
    }
    const long long sk_total = (tiles - dp_tiles) * KB;
    if (sk_total > 0 && sk == 0) { printf("FAIL: %lld stream-K k-blocks but no stream-K workers\n", sk_total); return 1; }
    std::vector<int> seen((size_t)tiles * KB, 0);
    long long most = 0;
    for (int w = 0; w < grid; ++w) {
        y2::CapIter it;
        it.init(w, grid, dp_tiles, sk, sk_total, KB, cap);
        int tile, kb0, kb1, a, b, prev_tile = -1, prev_kb1 = -1, prev_a = -1, seg_index = -1, first_len = 0;
        long long mine = 0;
        bool head_with_contrib_seen = false;
        while (it.next(tile, kb0, kb1, a, b)) {
            if (kb0 == a) { if (tile >= dp_tiles) ++seg_index; first_len = kb1 - kb0; }
            else if (tile != prev_tile || a != prev_a || kb0 != prev_kb1) { printf("FAIL: sub-segments of w=%d not contiguous\n", w); return 1; }
            if (head_with_contrib_seen && kb0 == a) { printf("FAIL: w=%d has a segment after a head that waits for contributors\n", w); return 1; }
            if (kb1 <= kb0 || kb0 < a || kb1 > b || b > KB || a < 0 || tile < 0 || tile >= tiles) { printf("FAIL: bad sub-segment\n"); return 1; }
            if (cap > 0 && kb1 - kb0 > cap) { printf("FAIL: chain of %d k-blocks exceeds the cap %d\n", kb1 - kb0, cap); return 1; }
            if (kb1 - kb0 > first_len) { printf("FAIL: unequal pieces\n"); return 1; }
            if (a != 0 && (tile < dp_tiles || seg_index != 0)) { printf("FAIL: w=%d: a partial (non-head) segment that is not the first of its stream-K phase\n", w); return 1; }
            if (a == 0 && b < KB && kb1 == b) head_with_contrib_seen = true;
            for (int k = kb0; k < kb1; ++k) ++seen[(size_t)tile * KB + k];
            mine += kb1 - kb0;
            prev_tile = tile; prev_kb1 = kb1; prev_a = a;
        }
        if (mine > most) most = mine;
    }
    for (size_t i = 0; i < seen.size(); ++i)
        if (seen[i] != 1) { printf("FAIL: tile %zu k-block %zu visited %d times (tiles=%lld KB=%d workers=%d cap=%d)\n", i / KB, i % KB, seen[i], tiles, KB, workers, cap); return 1; }
    if (verbose) printf("tiles=%lld KB=%d workers=%d cap=%d -> dp_tiles=%d sk_workers=%d grid=%d max_kblocks_per_worker=%lld\n", tiles, KB, workers, cap, dp_tiles, sk, grid, most);
    return 0;
}

int main() {
    int bad = 0;
    // the bench configuration (B=32, 416, C=80): single-CTA (148 workers) and CTA-pair (74 workers) plans of the 3x3 layers
    struct { long long tiles; int KB; size_t wbytes; } L[] = {{676, 18, 1179648}, {338, 36, 4718592}, {172, 72, 18874368}, {172, 144, 37748736},
                                                            {172, 432, 113246208}, {86, 16, 2097152}, {169, 8, 524288}};
    for (auto& l : L)
        for (int workers : {148, 74})
            for (int cap : {0, 16, 32, 36}) {
                const long long tiles = workers == 74 ? (l.tiles + 1) / 2 : l.tiles;
                bad += check(tiles, l.KB, workers, 0, 1.0, l.wbytes, cap, cap == 32);
            }
    // forced stream-K / capped grids (the test-suite variants) and random configurations
    for (int mc : {1, 2, -3, -5, -37, -74, 37}) bad += check(3, 72, 148, mc, 1.0, 0, 32, false);
    srand(7);
    for (int i = 0; i < 4000 && !bad; ++i) {
        const long long tiles = 1 + rand() % 700;
        const int KB = 1 + rand() % 450, workers = 1 + rand() % 148, cap = (rand() % 4 == 0) ? 0 : 1 + rand() % 64;
        const int mc = (rand() % 5 == 0) ? -(1 + rand() % workers) : ((rand() % 7 == 0) ? 1 + rand() % workers : 0);
        const size_t wb = (rand() % 2) ? (size_t)(rand() % 200) << 20 : 0;
        bad += check(tiles, KB, workers, mc, 0.25 + (rand() % 4) * 0.25, wb, cap, false);
    }
    printf(bad ? "SCHEDULE CHECK FAILED\n" : "SCHEDULE CHECK OK\n");
    return bad ? 1 : 0;
}
