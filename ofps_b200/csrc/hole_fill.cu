// MotionFieldDensifier::interpolate_empty_cells (ofps/src/motion_field.rs:193-294) — HOST code.
//
// The reference fills empty cells one at a time from a BTreeSet ordered by (-#filled 6-neighbours, index):
// every fill changes the keys of its neighbours and the value written depends on which neighbours were
// already filled, so the result is defined by a strictly sequential order (in the sparse case one region
// grows cell by cell from the first seed across the whole field).  There is no data-parallel form that keeps
// the reference's result, and its only caller writes .flo files offline (flow-extract/src/main.rs:81), so it
// stays on the host as sequential control logic, exactly as in the reference; the
// densifier in front of it and the per-pixel paths around it run on the GPU.  This is not a fallback of a
// GPU path: there is no GPU variant.
//
// The ordered set is restated as seven buckets (filled-neighbour count 0..6) of hierarchical bitmaps:
// "first element of the BTreeSet" = smallest index in the highest non-empty bucket; O(1)-ish per operation
// instead of O(log n) tree updates.
#include "common.cuh"

#include <cmath>

namespace ofpsb {

namespace {

class IndexSet {   // set of indices in [0, n): insert, erase, smallest element
public:
    explicit IndexSet(size_t n)
    {
        size_t words = (n + 63) / 64;
        for (;;) {
            levels_.emplace_back(words ? words : 1, 0ull);
            if (words <= 1) break;
            words = (words + 63) / 64;
        }
    }
    void insert(size_t i)
    {
        for (auto& lv : levels_) {
            const uint64_t before = lv[i >> 6];
            lv[i >> 6] = before | (1ull << (i & 63));
            if (before) break;   // the summary bits above are already set
            i >>= 6;
        }
    }
    void erase(size_t i)
    {
        for (auto& lv : levels_) {
            lv[i >> 6] &= ~(1ull << (i & 63));
            if (lv[i >> 6]) break;
            i >>= 6;
        }
    }
    bool first(size_t* out) const
    {
        if (!levels_.back()[0]) return false;
        size_t i = 0;
        for (size_t l = levels_.size(); l-- > 0;) i = (i << 6) | (size_t)__builtin_ctzll(levels_[l][i]);
        *out = i;
        return true;
    }

private:
    std::vector<std::vector<uint64_t>> levels_;
};

}  // namespace

void interpolate_empty_cells_host(float* sums, float* counts, size_t w, size_t h)
{
    const size_t cells = w * h;
    if (cells == 0) return;
    static const int NB[6][2] = {{-1, 0}, {0, -1}, {-1, -1}, {1, 0}, {0, 1}, {1, 1}};   // motion_field.rs:208
    auto filled_neighbours = [&](size_t i) {   // calc_counts (:210-228): neighbours with counts > 0.1
        const long x = (long)(i % w), y = (long)(i / w);
        int c = 0;
        for (const auto& o : NB) {
            const long nx = x + o[0], ny = y + o[1];
            if (nx >= 0 && nx < (long)w && ny >= 0 && ny < (long)h && counts[2 * ((size_t)nx + (size_t)ny * w)] > 0.1f) c++;
        }
        return c;
    };
    std::vector<int8_t> key(cells, -1);   // filled-neighbour count of a queued cell, -1 = not queued
    std::vector<IndexSet> bucket;
    bucket.reserve(7);
    for (int k = 0; k < 7; k++) bucket.emplace_back(cells);
    size_t queued = 0;
    for (size_t i = 0; i < cells; i++)
        if (counts[2 * i] < 0.5f) {   // :235
            key[i] = (int8_t)filled_neighbours(i);
            bucket[key[i]].insert(i);
            queued++;
        }
    if (queued == cells) return;   // no vectors at all (:243-245)
    while (queued) {
        size_t i = 0;
        int k = 6;
        while (k >= 0 && !bucket[k].first(&i)) k--;
        if (k <= 0) break;   // a cell without filled neighbour at the head: the reference would spin (:269-270); unreachable
        bucket[k].erase(i);
        key[i] = -1;
        queued--;
        const long x = (long)(i % w), y = (long)(i / w);
        for (const auto& o : NB) {   // :255-267
            const long nx = x + o[0], ny = y + o[1];
            if (nx < 0 || nx >= (long)w || ny < 0 || ny >= (long)h) continue;
            const size_t j = (size_t)nx + (size_t)ny * w;
            const float cnt = counts[2 * j];
            if (!(cnt > 0.1f)) continue;
            const float scale = 1.0f - std::sqrt((float)(o[0] * o[0] + o[1] * o[1])) * 0.5f;
            const float f = scale * (1.0f / cnt);                  // scale * inv_cnt, then scalar * column
            const float vx = f * sums[2 * j], vy = f * sums[2 * j + 1];
            counts[2 * i] += scale;                                // add_vector_idx (:141-147)
            counts[2 * i + 1] += scale;
            sums[2 * i] = vx * scale + sums[2 * i];
            sums[2 * i + 1] = vy * scale + sums[2 * i + 1];
        }
        for (const auto& o : NB) {   // :273-289: the newly filled cell raises the key of its queued neighbours
            const long nx = x + o[0], ny = y + o[1];
            if (nx < 0 || nx >= (long)w || ny < 0 || ny >= (long)h) continue;
            const size_t j = (size_t)nx + (size_t)ny * w;
            if (key[j] < 0) continue;
            bucket[key[j]].erase(j);
            key[j]++;
            bucket[key[j]].insert(j);
        }
    }
}

}  // namespace ofpsb
