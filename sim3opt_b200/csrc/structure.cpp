// structure.cpp -- sparsity pattern and index build.
//
// Replaces SparseOptimizer::initializeOptimization + BlockSolver::buildStructure [EXT g2o]
// (reference call site kitti_surf.cpp:674).  Free vertices are numbered in
// id order (fixed -> -1); Hpp holds upper-triangular blocks only:
//     {(k,k) for every free k}  U  {(min,max) of the Hessian indices of every edge with two free ends}.
// Internally blocks live in a BSR-upper array (row-major by block row, diagonal first), which is
// the order the vertex-pair-sorted edges arrive in; the g2o block-CCS view (column-major, rows
// ascending) is emitted alongside for the bit-exact structure check.
#include <algorithm>
#include <numeric>

#include "internal.h"

namespace s3o {

void build_structure_host(int nv, const uint8_t *fixed, int ne, const int32_t *v0, const int32_t *v1,
                          HostStructure &S) {
    std::vector<int32_t> hidx(nv);
    int nf = 0;
    for (int v = 0; v < nv; ++v) hidx[v] = (fixed && fixed[v]) ? -1 : nf++;
    build_structure_from_hidx(nv, hidx.data(), nf, ne, v0, v1, S);
}

void build_structure_from_hidx(int nv, const int32_t *hidx_in, int nfree, int ne, const int32_t *v0,
                               const int32_t *v1, HostStructure &S) {
    S = HostStructure();
    S.nv = nv;
    S.ne = ne;
    S.hidx.assign(hidx_in, hidx_in + nv);
    S.free2v.assign(nfree, -1);
    for (int v = 0; v < nv; ++v)
        if (S.hidx[v] >= 0) S.free2v[S.hidx[v]] = v;
    const int nf = S.nf = nfree;

    // sort active edges by (row=min, col=max) Hessian pair; one-free-end edges sort at (h,h).
    // Ties keep the caller's edge order, so the summation order inside a block is reproducible.
    std::vector<uint64_t> key;
    key.reserve(ne);
    S.perm.clear();
    S.perm.reserve(ne);
    for (int k = 0; k < ne; ++k) {
        const int hi = S.hidx[v0[k]], hj = S.hidx[v1[k]];
        if (hi < 0 && hj < 0) continue;  // g2o drops edges whose vertices are all fixed
        S.perm.push_back(k);
    }
    const int na = S.ne_act = (int)S.perm.size();
    key.resize(ne);
    for (int k = 0; k < ne; ++k) {
        const int hi = S.hidx[v0[k]], hj = S.hidx[v1[k]];
        int r, c;
        if (hi < 0) r = c = hj;
        else if (hj < 0) r = c = hi;
        else { r = std::min(hi, hj); c = std::max(hi, hj); }
        key[k] = ((uint64_t)(uint32_t)r << 32) | (uint32_t)c;
    }
    std::stable_sort(S.perm.begin(), S.perm.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });

    S.sv0.resize(na);
    S.sv1.resize(na);
    S.e_blk.assign(na, -1);
    // BSR upper: diagonal first, then distinct off-diagonal columns ascending
    S.rowptr.assign(nf + 1, 0);
    S.colidx.clear();
    S.colidx.reserve(nf + na);
    S.blk_ebeg.clear();
    S.blk_eend.clear();
    S.blk_ebeg.reserve(nf + na);
    S.blk_eend.reserve(nf + na);
    int s = 0;
    for (int r = 0; r < nf; ++r) {
        S.rowptr[r] = (int32_t)S.colidx.size();
        S.colidx.push_back(r);
        S.blk_ebeg.push_back(0);  // diagonal block: filled through the incidence lists, no edge range
        S.blk_eend.push_back(0);
        uint64_t last = ~0ull;
        while (s < na && (int)(key[S.perm[s]] >> 32) == r) {
            const uint64_t kk = key[S.perm[s]];
            const int c = (int)(uint32_t)kk;
            if (c != r) {
                if (kk != last) {
                    S.colidx.push_back(c);
                    S.blk_ebeg.push_back(s);
                    S.blk_eend.push_back(s);
                    last = kk;
                }
                S.e_blk[s] = (int32_t)S.colidx.size() - 1;
                S.blk_eend.back() = s + 1;
            }
            ++s;
        }
    }
    S.rowptr[nf] = (int32_t)S.colidx.size();
    const int nb = S.nb = (int)S.colidx.size();
    for (int t = 0; t < na; ++t) {
        S.sv0[t] = v0[S.perm[t]];
        S.sv1[t] = v1[S.perm[t]];
    }

    // off-diagonal blocks fed by a single edge (the common case) carry their source directly
    S.blk_src.assign(nb, -1);
    for (int k = 0; k < nb; ++k)
        if (S.blk_eend[k] - S.blk_ebeg[k] == 1) {
            const int t = S.blk_ebeg[k];
            const bool transposed = S.hidx[S.sv0[t]] > S.hidx[S.sv1[t]];   // vertex(0) on the max side: A^T O' B is the transpose
            S.blk_src[k] = (t << 1) | (transposed ? 1 : 0);
        } else if (S.blk_eend[k] - S.blk_ebeg[k] > 1) {
            S.multi_blk.push_back(k);
        }

    // incidences per free vertex, ordered by sorted edge position (fixed summation order)
    S.inc_ptr.assign(nf + 1, 0);
    for (int t = 0; t < na; ++t) {
        const int hi = S.hidx[S.sv0[t]], hj = S.hidx[S.sv1[t]];
        if (hi >= 0) S.inc_ptr[hi + 1]++;
        if (hj >= 0) S.inc_ptr[hj + 1]++;
    }
    for (int r = 0; r < nf; ++r) S.inc_ptr[r + 1] += S.inc_ptr[r];
    S.inc_ent.resize(S.inc_ptr[nf]);
    {
        std::vector<int32_t> fill(S.inc_ptr.begin(), S.inc_ptr.end() - 1);
        for (int t = 0; t < na; ++t) {
            const int hi = S.hidx[S.sv0[t]], hj = S.hidx[S.sv1[t]];
            if (hi >= 0) S.inc_ent[fill[hi]++] = (t << 1);
            if (hj >= 0) S.inc_ent[fill[hj]++] = (t << 1) | 1;
        }
    }

    // column view of the off-diagonal blocks (for the transposed half of the symmetric SpMV)
    S.colT_ptr.assign(nf + 1, 0);
    for (int r = 0; r < nf; ++r)
        for (int k = S.rowptr[r] + 1; k < S.rowptr[r + 1]; ++k) S.colT_ptr[S.colidx[k] + 1]++;
    for (int c = 0; c < nf; ++c) S.colT_ptr[c + 1] += S.colT_ptr[c];
    S.colT_blk.resize(S.colT_ptr[nf]);
    {
        std::vector<int32_t> fill(S.colT_ptr.begin(), S.colT_ptr.end() - 1);
        for (int r = 0; r < nf; ++r)  // rows ascending => each column list is ordered by row
            for (int k = S.rowptr[r] + 1; k < S.rowptr[r + 1]; ++k) S.colT_blk[fill[S.colidx[k]]++] = k;
    }

    // g2o-order upper block-CCS: column c holds rows r<=c ascending, diagonal last
    S.ccs_colptr.assign(nf + 1, 0);
    S.ccs_rowidx.resize(nb);
    S.ccs2bsr.resize(nb);
    int pos = 0;
    for (int c = 0; c < nf; ++c) {
        S.ccs_colptr[c] = pos;
        for (int t = S.colT_ptr[c]; t < S.colT_ptr[c + 1]; ++t) {
            const int k = S.colT_blk[t];
            // row of block k: binary search in rowptr
            const int r = (int)(std::upper_bound(S.rowptr.begin(), S.rowptr.end(), k) - S.rowptr.begin()) - 1;
            S.ccs_rowidx[pos] = r;
            S.ccs2bsr[pos] = k;
            ++pos;
        }
        S.ccs_rowidx[pos] = c;
        S.ccs2bsr[pos] = S.rowptr[c];
        ++pos;
    }
    S.ccs_colptr[nf] = pos;
}


void build_tiles(const std::vector<int32_t> &rowptr, int nf, int cap, std::vector<int32_t> &tile_row) {
    tile_row.clear();
    tile_row.push_back(0);
    int r = 0;
    while (r < nf) {
        const int start = r;
        int blocks = 0;
        while (r < nf) {
            const int nb_row = rowptr[r + 1] - rowptr[r];
            if (blocks + nb_row > cap && r > start) break;
            blocks += nb_row;
            ++r;
            if (blocks >= cap) break;   // also closes a hub row's own tile
        }
        tile_row.push_back(r);
    }
}

void build_partition_plan(int nv, const uint8_t *fixed, int ne, const int32_t *v0, const int32_t *v1, int rank,
                          int world, PartitionPlan &P) {
    P = PartitionPlan();
    P.rank = rank;
    P.world = world;
    P.ghidx.resize(nv);
    int nf = 0;
    for (int v = 0; v < nv; ++v) P.ghidx[v] = (fixed && fixed[v]) ? -1 : nf++;
    P.nf_global = nf;
    P.seg = world > 0 ? (nf + world - 1) / world : nf;
    if (P.seg == 0) P.seg = 1;
    P.own_lo = std::min(nf, rank * P.seg);
    const int own_hi = std::min(nf, (rank + 1) * P.seg);
    P.n_own = own_hi - P.own_lo;
    auto owner = [&](int g) { return g / P.seg; };
    // local edges, ghosts and what every peer needs from us
    std::vector<std::vector<int32_t>> need_from_me(world);   // global Hessian indices I own that peer q needs
    for (int k = 0; k < ne; ++k) {
        const int ga = P.ghidx[v0[k]], gb = P.ghidx[v1[k]];
        if (ga < 0 && gb < 0) continue;
        const int oa = ga >= 0 ? owner(ga) : -1, ob = gb >= 0 ? owner(gb) : -1;
        if (oa != rank && ob != rank) continue;
        P.local_edges.push_back(k);
        int gmin = ga < 0 ? gb : (gb < 0 ? ga : std::min(ga, gb));
        P.primary.push_back(owner(gmin) == rank ? 1 : 0);
        if (ga >= 0 && gb >= 0 && oa != ob) {
            if (oa == rank) { P.ghosts.push_back(gb); need_from_me[ob].push_back(ga); }
            else            { P.ghosts.push_back(ga); need_from_me[oa].push_back(gb); }
        }
    }
    std::sort(P.ghosts.begin(), P.ghosts.end());
    P.ghosts.erase(std::unique(P.ghosts.begin(), P.ghosts.end()), P.ghosts.end());
    P.n_ghost = (int)P.ghosts.size();
    P.recv_count.assign(world, 0);
    P.recv_off.assign(world, 0);
    for (int g : P.ghosts) P.recv_count[owner(g)]++;
    for (int q = 1; q < world; ++q) P.recv_off[q] = P.recv_off[q - 1] + P.recv_count[q - 1];
    P.send_count.assign(world, 0);
    P.send_off.assign(world, 0);
    P.send_idx.clear();
    for (int q = 0; q < world; ++q) {
        auto &lst = need_from_me[q];
        std::sort(lst.begin(), lst.end());
        lst.erase(std::unique(lst.begin(), lst.end()), lst.end());
        P.send_off[q] = (int32_t)P.send_idx.size();
        P.send_count[q] = (int32_t)lst.size();
        for (int g : lst) P.send_idx.push_back(g - P.own_lo);
    }
    // local Hessian numbering: owned first (global order), then ghosts (global order)
    P.lhidx.assign(nv, -1);
    std::vector<int32_t> g2l(nf, -1);
    for (int g = P.own_lo; g < own_hi; ++g) g2l[g] = g - P.own_lo;
    for (int t = 0; t < P.n_ghost; ++t) g2l[P.ghosts[t]] = P.n_own + t;
    for (int v = 0; v < nv; ++v)
        if (P.ghidx[v] >= 0) P.lhidx[v] = g2l[P.ghidx[v]];
}

}  // namespace s3o
