// Host-only (plain C++17, no CUDA): the splitters of a prefix-sharded set.  Included by sharded_index.cu; compiled on its own
// by tests/test_sharded_gloo.py, which checks it against cbl_b200/sharded.py's equal_mass_splitters on the same sample.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cbl {

// world - 1 splitters s.t. each range holds ~1/world of the sample's COST (the owner-side probe time, not the word count).
// Relative cost of a word as a piecewise-linear function of the mass quantile of its prefix in the sorted sample: flat but
// for the sparse tail of the prefix space (measured on B200 against the shards of an 8-GPU set, scripts/exp_shard_probe.py;
// the same knots as PROBE_COST_KNOTS in cbl_b200/sharded.py)
inline std::vector<uint32_t> equal_cost_splitters(std::vector<uint32_t> pre, int world) {
    static const double KQ[] = {0.0, 0.625, 0.875, 0.9375, 1.0}, KW[] = {1.0, 1.0, 1.05, 1.12, 1.2};
    std::vector<uint32_t> sp;
    if (world <= 1 || pre.empty()) return sp;
    std::sort(pre.begin(), pre.end());
    const size_t n = pre.size();
    const int GRID = 4096;
    auto weight = [&](double q) {
        int k = 0;
        while (k < 3 && q > KQ[k + 1]) k++;
        return KW[k] + (KW[k + 1] - KW[k]) * (q - KQ[k]) / (KQ[k + 1] - KQ[k]);
    };
    std::vector<double> cum(GRID + 1, 0.0);   // cumulative cost C(q) on a grid
    for (int i = 1; i <= GRID; i++) cum[i] = cum[i - 1] + 0.5 * (weight((double)(i - 1) / GRID) + weight((double)i / GRID)) / GRID;
    for (int i = 1; i < world; i++) {
        const double c = cum[GRID] * i / world;
        const int hi = (int)(std::upper_bound(cum.begin(), cum.end(), c) - cum.begin());   // cum[hi - 1] <= c < cum[hi]
        const int lo = std::max(hi - 1, 0);
        const double f = hi <= GRID && cum[hi] > cum[lo] ? (c - cum[lo]) / (cum[hi] - cum[lo]) : 0.0;
        const double q = ((double)lo + f) / GRID;
        sp.push_back(pre[std::min<size_t>((size_t)(q * n), n - 1)]);
    }
    for (size_t i = 1; i < sp.size(); i++) if (sp[i] <= sp[i - 1]) sp[i] = sp[i - 1] + 1;   // strictly increasing
    return sp;
}


}  // namespace cbl
