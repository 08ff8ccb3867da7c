// Linear sum assignment (square, K <= 32) as one sequential routine usable on host and device.
//
// Replaces scipy.optimize.linear_sum_assignment(iou, maximize=True) in match_mask_by_iou
// (losses/seg_loss_unsup.py:234-237; scipy is an un-vendored dependency of the reference, pinned only as
// "scipy" in requirements.txt; installed here: 1.18.1).  scipy implements the shortest-augmenting-path
// algorithm of D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016;
// this is a restatement of that published algorithm INCLUDING its tie-breaking (columns scanned through the
// `remaining` list that starts in descending order; among equal reduced costs an unassigned column wins),
// because empty slots make all-zero IoU rows -- ties are the common case.  The port is checked against scipy
// on tie-heavy matrices in tests/test_lsap.py (host twin) and tests/test_gpu_step.py::test_device_hungarian_matches_scipy (device).
#pragma once
#include <math.h>

namespace ogc {

constexpr int kLsapMax = 32;

// cost: n x n row-major (already negated for maximisation); col4row[i] = column assigned to row i
__host__ __device__ inline void lsap_solve(int n, const double *cost, int *col4row) {
    double u[kLsapMax], v[kLsapMax], spc[kLsapMax];
    int path[kLsapMax], row4col[kLsapMax], remaining[kLsapMax];
    bool SR[kLsapMax], SC[kLsapMax];
    for (int i = 0; i < n; ++i) { u[i] = 0.0; v[i] = 0.0; path[i] = -1; col4row[i] = -1; row4col[i] = -1; }
    for (int cur = 0; cur < n; ++cur) {
        // ---- shortest augmenting path from row `cur` ----
        double minVal = 0.0;
        int num_remaining = n;
        for (int it = 0; it < n; ++it) { remaining[it] = n - it - 1; SR[it] = false; SC[it] = false; spc[it] = INFINITY; }
        int sink = -1, i = cur;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = true;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[it];
                const double r = minVal + cost[i * n + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            minVal = lowest;
            if (!(minVal < INFINITY)) { sink = -2; break; }   // infeasible (cannot happen for finite costs)
            const int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = true;
            remaining[index] = remaining[--num_remaining];
        }
        if (sink < 0) return;
        // ---- dual update ----
        u[cur] += minVal;
        for (int r = 0; r < n; ++r)
            if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < n; ++j)
            if (SC[j]) v[j] -= minVal - spc[j];
        // ---- augment ----
        int j = sink;
        while (true) {
            const int r = path[j];
            row4col[j] = r;
            const int prev = col4row[r];
            col4row[r] = j;
            j = prev;
            if (r == cur) break;
        }
    }
}

}  // namespace ogc
