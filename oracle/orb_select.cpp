// oracle/orb_select.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.py).
// Restates KeyPointsFilter::retainBest of OpenCV 4.13.0 (features2d, keypoint.cpp): std::nth_element on the response (greater-than
// order), then std::partition of the tail on "response >= the n-th response", so ties of the boundary response are all kept.  The
// ORDER the survivors come out in is whatever libstdc++'s introselect leaves behind -- and it is observable, because the row order of
// a frame's descriptors decides the lowest-index tie-breaks of the matcher -- so the oracle calls the same two library functions on
// the same sequence rather than re-deriving their element moves.
#include <algorithm>
#include <cstdint>
#include <vector>

namespace {
struct Item { float response; int32_t index; };
}

// responses[n] in detection order; writes the surviving original indices to out_index (capacity n) in the order retainBest leaves
// them; returns how many.
extern "C" int orb_oracle_retain_best(const float* responses, int n, int n_points, int32_t* out_index) {
    std::vector<Item> v(n);
    for (int i = 0; i < n; ++i) v[i] = Item{responses[i], i};
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const Item& a, const Item& b) { return a.response > b.response; });
        const float ambiguous = v[n_points - 1].response;
        auto new_end = std::partition(v.begin() + n_points, v.end(), [ambiguous](const Item& a) { return a.response >= ambiguous; });
        v.resize(new_end - v.begin());
    }
    for (size_t i = 0; i < v.size(); ++i) out_index[i] = v[i].index;
    return (int)v.size();
}
