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
    S = HostStructure();
    S.nv = nv;
    S.ne = ne;
    S.hidx.resize(nv);
    S.free2v.clear();
    for (int v = 0; v < nv; ++v) {
        if (fixed && fixed[v]) S.hidx[v] = -1;
        else { S.hidx[v] = (int32_t)S.free2v.size(); S.free2v.push_back(v); }
    }
    const int nf = S.nf = (int)S.free2v.size();

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

}  // namespace s3o
