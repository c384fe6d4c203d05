// amg_host.cpp -- aggregation hierarchy of the multilevel preconditioner (structure only, host).
// See amg.h.  Deterministic: vertices are visited in index order, lists are sorted.
#include <algorithm>
#include <thread>

#include "amg.h"

namespace s3o {
namespace {

// LSD radix sort of 64-bit keys on their low `bits` bits (11 bits per pass): the coarse-pattern keys of a 1M-pose
// graph are 12 M entries, where std::sort costs most of a second
void radix_sort_u64(std::vector<uint64_t> &a, int bits) {
    if (a.size() < 4096) { std::sort(a.begin(), a.end()); return; }
    constexpr int R = 11, B = 1 << R;
    std::vector<uint64_t> tmp(a.size());
    std::vector<size_t> cnt(B);
    uint64_t *src = a.data(), *dst = tmp.data();
    for (int shift = 0; shift < bits; shift += R) {
        std::fill(cnt.begin(), cnt.end(), 0);
        for (size_t t = 0; t < a.size(); ++t) cnt[(src[t] >> shift) & (B - 1)]++;
        size_t run = 0;
        for (int b = 0; b < B; ++b) { const size_t c = cnt[b]; cnt[b] = run; run += c; }
        for (size_t t = 0; t < a.size(); ++t) dst[cnt[(src[t] >> shift) & (B - 1)]++] = src[t];
        std::swap(src, dst);
    }
    if (src != a.data()) std::copy(src, src + a.size(), a.data());
}

// fn(lo, hi) over [0, n) on a few host threads (disjoint index ranges: the result does not depend on the count)
template <class F>
void parallel_ranges(int n, F fn) {
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (n < 100000 || hw == 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (unsigned w = 0; w < hw; ++w) th.emplace_back([=]() { fn((int)((long long)n * w / hw), (int)((long long)n * (w + 1) / hw)); });
    for (auto &t : th) t.join();
}

struct Pattern {            // blocks of one level
    int n = 0, nblk = 0;
    const int32_t *brow = nullptr, *bcol = nullptr;
    bool upper = false;     // level 0: only (i <= j) stored
};

// adjacency (both directions, self excluded), neighbours ascending; seg > 0: pairs whose ends lie in
// different index segments (ranks of the partitioned solve) are left out, so no aggregate crosses a cut
void build_adjacency(const Pattern &F, int seg, std::vector<int32_t> &ptr, std::vector<int32_t> &idx) {
    ptr.assign(F.n + 1, 0);
    for (int k = 0; k < F.nblk; ++k) {
        const int i = F.brow[k], j = F.bcol[k];
        if (i == j || (seg > 0 && i / seg != j / seg)) continue;
        ptr[i + 1]++;
        if (F.upper) ptr[j + 1]++;
    }
    for (int i = 0; i < F.n; ++i) ptr[i + 1] += ptr[i];
    idx.resize(ptr[F.n]);
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int k = 0; k < F.nblk; ++k) {
        const int i = F.brow[k], j = F.bcol[k];
        if (i == j || (seg > 0 && i / seg != j / seg)) continue;
        idx[fill[i]++] = j;
        if (F.upper) idx[fill[j]++] = i;
    }
    parallel_ranges(F.n, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) std::sort(idx.begin() + ptr[i], idx.begin() + ptr[i + 1]);
    });
}

// Greedy neighbourhood aggregation: a vertex whose whole neighbourhood is still free seeds an
// aggregate {vertex + neighbours}; leftovers join the aggregate of their first aggregated
// neighbour, isolated leftovers stay alone.
void aggregate(int n, const std::vector<int32_t> &ptr, const std::vector<int32_t> &idx, std::vector<int32_t> &agg,
               std::vector<int32_t> &root) {
    agg.assign(n, -1);
    root.clear();
    for (int i = 0; i < n; ++i) {
        if (agg[i] >= 0) continue;
        bool free_nb = true;
        for (int t = ptr[i]; t < ptr[i + 1] && free_nb; ++t) free_nb = agg[idx[t]] < 0;
        if (!free_nb) continue;
        const int a = (int)root.size();
        root.push_back(i);
        agg[i] = a;
        for (int t = ptr[i]; t < ptr[i + 1]; ++t) agg[idx[t]] = a;
    }
    const int n_seeded = (int)root.size();
    std::vector<int32_t> join(n, -1);
    for (int i = 0; i < n; ++i) {
        if (agg[i] >= 0) continue;
        for (int t = ptr[i]; t < ptr[i + 1]; ++t) {
            const int a = agg[idx[t]];
            if (a >= 0 && a < n_seeded) { join[i] = a; break; }
        }
    }
    for (int i = 0; i < n; ++i) {
        if (agg[i] >= 0) continue;
        if (join[i] >= 0) agg[i] = join[i];
        else { agg[i] = (int)root.size(); root.push_back(i); }
    }
    // number the aggregates by ascending root, so coarse indices follow the fine order (and a
    // vertex-range partition of the fine level induces contiguous coarse ranges)
    const int na = (int)root.size();
    std::vector<int32_t> order(na), newid(na);
    for (int a = 0; a < na; ++a) order[a] = a;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return root[a] < root[b]; });
    std::vector<int32_t> sorted_root(na);
    for (int t = 0; t < na; ++t) { newid[order[t]] = t; sorted_root[t] = root[order[t]]; }
    root.swap(sorted_root);
    for (int i = 0; i < n; ++i) agg[i] = newid[agg[i]];
}

int find_col(const AmgHostLevel &L, int I, int J) {
    const int32_t *b = L.colidx.data() + L.rowptr[I], *e = L.colidx.data() + L.rowptr[I + 1];
    return (int)(std::lower_bound(b, e, J) - L.colidx.data());
}

}  // namespace

void amg_build_hierarchy(const HostStructure &S, int coarsest_max, int max_levels, std::vector<AmgHostLevel> &levels,
                         int seg) {
    levels.clear();
    if (S.nf == 0) return;
    levels.reserve((size_t)max_levels + 1);     // the loop keeps pointers into levels.back()
    std::vector<int32_t> blk_row0(S.nb);
    for (int r = 0; r < S.nf; ++r)
        for (int k = S.rowptr[r]; k < S.rowptr[r + 1]; ++k) blk_row0[k] = r;
    Pattern F;
    F.n = S.nf; F.nblk = S.nb; F.brow = blk_row0.data(); F.bcol = S.colidx.data(); F.upper = true;
    const std::vector<int32_t> *vid_fine = &S.free2v;

    while (F.n > coarsest_max && (int)levels.size() < max_levels) {
        std::vector<int32_t> aptr, aidx;
        build_adjacency(F, levels.empty() ? seg : 0, aptr, aidx);
        setup_mark("    adjacency");
        AmgHostLevel L;
        L.n_fine = F.n;
        aggregate(F.n, aptr, aidx, L.agg, L.root);
        setup_mark("    aggregate");
        L.n = (int)L.root.size();
        if (L.n * 10 > F.n * 9) break;          // aggregation stalled (isolated vertices): stop here
        L.vid.resize(L.n);
        for (int a = 0; a < L.n; ++a) L.vid[a] = (*vid_fine)[L.root[a]];
        // members
        L.mem_ptr.assign(L.n + 1, 0);
        for (int i = 0; i < F.n; ++i) L.mem_ptr[L.agg[i] + 1]++;
        for (int a = 0; a < L.n; ++a) L.mem_ptr[a + 1] += L.mem_ptr[a];
        L.mem_idx.resize(F.n);
        {
            std::vector<int32_t> fill(L.mem_ptr.begin(), L.mem_ptr.end() - 1);
            for (int i = 0; i < F.n; ++i) L.mem_idx[fill[L.agg[i]]++] = i;
        }
        // coarse pattern: keys (I << bits) | J of every fine block (both orientations for an upper-only level), sorted and
        // de-duplicated.  Fine blocks of neighbouring rows fall into the same few coarse blocks, so each host thread first
        // sorts and de-duplicates its own contiguous range (12 M keys shrink to ~1 M), then one sort merges the ranges;
        // the result is the sorted set of distinct keys whatever the thread count.
        std::vector<uint64_t> keys;
        {
            int bits = 1;
            while ((1ll << bits) <= L.n) ++bits;
            const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
            const unsigned parts = F.nblk < 100000 ? 1u : hw;
            std::vector<std::vector<uint64_t>> part(parts);
            auto work = [&](unsigned w) {
                const int lo = (int)((long long)F.nblk * w / parts), hi = (int)((long long)F.nblk * (w + 1) / parts);
                std::vector<uint64_t> &K = part[w];
                K.reserve((size_t)(hi - lo) * (F.upper ? 2 : 1));
                for (int k = lo; k < hi; ++k) {
                    const uint64_t I = (uint32_t)L.agg[F.brow[k]], J = (uint32_t)L.agg[F.bcol[k]];
                    K.push_back((I << bits) | J);
                    if (F.upper && I != J) K.push_back((J << bits) | I);
                }
                radix_sort_u64(K, 2 * bits);
                K.erase(std::unique(K.begin(), K.end()), K.end());
            };
            if (parts == 1) work(0);
            else {
                std::vector<std::thread> th;
                for (unsigned w = 0; w < parts; ++w) th.emplace_back(work, w);
                for (auto &t : th) t.join();
            }
            size_t total = 0;
            for (auto &K : part) total += K.size();
            keys.reserve(total);
            for (auto &K : part) keys.insert(keys.end(), K.begin(), K.end());
            if (parts > 1) {
                radix_sort_u64(keys, 2 * bits);
                keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
            }
            for (uint64_t &k : keys) k = ((k >> bits) << 32) | (k & ((1ull << bits) - 1));
        }
        setup_mark("    coarse keys sort");
        const int nblk = (int)keys.size();
        L.rowptr.assign(L.n + 1, 0);
        L.colidx.resize(nblk);
        L.blk_row.resize(nblk);
        L.dpos.assign(L.n, -1);
        for (int t = 0; t < nblk; ++t) {
            const int I = (int)(keys[t] >> 32), J = (int)(uint32_t)keys[t];
            L.rowptr[I + 1]++;
            L.colidx[t] = J;
            L.blk_row[t] = I;
            if (I == J) L.dpos[I] = t;
        }
        for (int a = 0; a < L.n; ++a) L.rowptr[a + 1] += L.rowptr[a];
        // upper blocks and their mirrors
        std::vector<int32_t> ubidx(nblk, -1);
        for (int t = 0; t < nblk; ++t) {
            const int I = L.blk_row[t], J = L.colidx[t];
            if (I > J) continue;
            ubidx[t] = (int32_t)L.gal_out.size();
            L.gal_out.push_back(t);
            L.gal_mirror.push_back(I == J ? -1 : find_col(L, J, I));
            L.gal_I.push_back(I);
            L.gal_J.push_back(J);
        }
        L.nub = (int)L.gal_out.size();
        // contributors, ascending in the fine block index
        auto target = [&](int k, int &ub, int &flag) {
            const int i = F.brow[k], j = F.bcol[k];
            const int I = L.agg[i], J = L.agg[j];
            if (F.upper) {
                if (i == j) { ub = ubidx[L.dpos[I]]; flag = 0; }
                else if (I == J) { ub = ubidx[L.dpos[I]]; flag = 2; }
                else if (I < J) { ub = ubidx[find_col(L, I, J)]; flag = 0; }
                else { ub = ubidx[find_col(L, J, I)]; flag = 1; }
            } else {
                if (I > J) { ub = -1; flag = 0; }
                else { ub = ubidx[find_col(L, I, J)]; flag = 0; }
            }
        };
        setup_mark("    coarse pattern");
        L.gal_ptr.assign(L.nub + 1, 0);
        std::vector<int32_t> tgt(F.nblk);            // (ub << 2) | flag or -1, computed once (binary searches)
        parallel_ranges(F.nblk, [&](int lo, int hi) {
            for (int k = lo; k < hi; ++k) {
                int ub, flag;
                target(k, ub, flag);
                tgt[k] = ub >= 0 ? (ub << 2) | flag : -1;
            }
        });
        setup_mark("    targets");
        for (int k = 0; k < F.nblk; ++k)
            if (tgt[k] >= 0) L.gal_ptr[(tgt[k] >> 2) + 1]++;
        for (int u = 0; u < L.nub; ++u) L.gal_ptr[u + 1] += L.gal_ptr[u];
        L.gal_ent.resize(L.gal_ptr[L.nub]);
        L.gal_i.resize(L.gal_ent.size());
        L.gal_j.resize(L.gal_ent.size());
        {
            std::vector<int32_t> fill(L.gal_ptr.begin(), L.gal_ptr.end() - 1);
            for (int k = 0; k < F.nblk; ++k) {
                if (tgt[k] < 0) continue;
                const int ub = tgt[k] >> 2, flag = tgt[k] & 3;
                const int pos = fill[ub]++;
                L.gal_ent[pos] = (k << 2) | flag;
                L.gal_i[pos] = F.brow[k];
                L.gal_j[pos] = F.bcol[k];
            }
        }
        setup_mark("    contributor lists");
        levels.push_back(std::move(L));
        const AmgHostLevel &B = levels.back();
        F.n = B.n; F.nblk = (int)B.colidx.size(); F.brow = B.blk_row.data(); F.bcol = B.colidx.data(); F.upper = false;
        vid_fine = &B.vid;
    }
}

}  // namespace s3o
