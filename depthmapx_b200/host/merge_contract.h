// Merge links for the BFS analyses: contraction of merged pairs and the at-the-radius correction (SURVEY.md §8 row f3,
// A.3).  Plain functions over CSR arrays so that both the stand-alone host layer (dmx::PointMap::contractedRows /
// radiusCorrection) and the shim translation units inside the reference tree (integration/) use the same code.
//
// Reference behaviour (vgavisualglobal.cpp:108-121, vgavisualglobaldepth.cpp:54-63): a cell that is expanded also expands
// its merge partner and finalises it without counting it, i.e. a merged pair behaves as ONE vertex reached at the smaller
// of the two levels and counted once.  The pair is contracted into its smaller-ordinal cell (`primary`): that row becomes
// the union of both rows, every edge into either cell points at it, the other cell keeps an empty row; results of the
// secondary are copies of the primary's.  With a radius R the one difference is that cells AT the radius are counted but
// not expanded, so a pair whose two cells are BOTH first reached at exactly level R (each through an edge to that very
// cell) counts twice in the reference: radius_correction adds that second count.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

namespace dmx {

struct Contracted {
    int64_t n = 0, ghosts = 0;
    std::vector<uint64_t> rowptr;
    std::vector<uint32_t> col;
    std::vector<int32_t> primary;       // [n] ordinal whose results a cell takes
    std::vector<int32_t> merged_cells;  // ordinals of all merged cells, pairs adjacent: (primary, secondary)*
    std::vector<uint64_t> in_ptr;       // [merged_cells.size()+1]
    std::vector<int32_t> in_list;       // contracted vertices whose (union) row holds the cell itself
};

// BFS "level of every vertex from a set of seeds" over a CSR handed over once (the transposed contracted adjacency):
// vga_step_depth on the GPU in the product, any BFS in the CPU tests.
struct LevelTo {
    virtual ~LevelTo() {}
    virtual void prepare(int64_t n, const std::vector<uint64_t> &t_rowptr, const std::vector<uint32_t> &t_col) = 0;
    virtual void run(const std::vector<int64_t> &seeds, std::vector<int32_t> &level) = 0;
};

// rows: ordinals (< n cells, >= n ghosts); partner[v] = ordinal of the merge partner or -1 (symmetric, exclusive)
inline void contract_rows(int64_t n, const uint64_t *rowptr, const uint32_t *col, const int32_t *partner, Contracted &out) {
    out.n = n;
    out.primary.resize((size_t)n);
    for (int64_t v = 0; v < n; v++) out.primary[(size_t)v] = partner[v] >= 0 ? std::min<int32_t>((int32_t)v, partner[v]) : (int32_t)v;
    // for every merged cell: the contracted vertices with an edge to the cell itself (radius_correction)
    out.merged_cells.clear();
    std::vector<int32_t> slot((size_t)n, -1);
    for (int64_t v = 0; v < n; v++)
        if (partner[v] > v) {
            slot[(size_t)v] = (int32_t)out.merged_cells.size();
            out.merged_cells.push_back((int32_t)v);
            slot[(size_t)partner[v]] = (int32_t)out.merged_cells.size();
            out.merged_cells.push_back(partner[v]);
        }
    std::vector<std::vector<int32_t>> ins(out.merged_cells.size());
    for (int64_t u = 0; u < n; u++)
        for (uint64_t e = rowptr[u]; e < rowptr[u + 1]; e++)
            if (col[e] < (uint32_t)n && slot[col[e]] >= 0) ins[(size_t)slot[col[e]]].push_back(out.primary[(size_t)u]);
    out.in_ptr.assign(1, 0);
    out.in_list.clear();
    for (auto &l : ins) {
        std::sort(l.begin(), l.end());
        l.erase(std::unique(l.begin(), l.end()), l.end());
        out.in_list.insert(out.in_list.end(), l.begin(), l.end());
        out.in_ptr.push_back(out.in_list.size());
    }
    // contracted rows: redirect every column to its primary, union the pair's rows into the primary, sort + unique
    out.rowptr.assign(1, 0);
    out.col.clear();
    out.col.reserve((size_t)rowptr[n]);
    std::vector<uint32_t> row;
    for (int64_t v = 0; v < n; v++) {
        row.clear();
        if (out.primary[(size_t)v] == v) {
            const int64_t members[2] = {v, partner[v]};
            for (int64_t u : members) {
                if (u < 0) continue;
                for (uint64_t e = rowptr[u]; e < rowptr[u + 1]; e++) {
                    const uint32_t c = col[e];
                    row.push_back(c < (uint32_t)n ? (uint32_t)out.primary[c] : c);
                }
            }
            std::sort(row.begin(), row.end());
            row.erase(std::unique(row.begin(), row.end()), row.end());
        }
        out.col.insert(out.col.end(), row.begin(), row.end());
        out.rowptr.push_back(out.col.size());
    }
}

// adds to the integers of the contracted BFS (all n sources) the second count of every pair whose cells are both first
// reached at level `radius`: total_nodes += 1, total_depth += radius, dist[radius] += 1
inline void radius_correction(const Contracted &c, int radius, LevelTo &level_to, int32_t *total_nodes, int64_t *total_depth,
                              int32_t *dist, int32_t max_levels) {
    if (radius < 1 || radius >= max_levels || c.merged_cells.empty()) return;
    const int64_t n = c.n;
    // transpose of the contracted adjacency (cells only: ghosts are never expanded)
    std::vector<uint64_t> t_rowptr((size_t)n + 1, 0);
    for (uint32_t x : c.col)
        if (x < (uint32_t)n) t_rowptr[(size_t)x + 1]++;
    for (int64_t v = 0; v < n; v++) t_rowptr[(size_t)v + 1] += t_rowptr[(size_t)v];
    std::vector<uint32_t> t_col(t_rowptr[(size_t)n]);
    std::vector<uint64_t> fill(t_rowptr.begin(), t_rowptr.end() - 1);
    for (int64_t u = 0; u < n; u++)
        for (uint64_t e = c.rowptr[(size_t)u]; e < c.rowptr[(size_t)u + 1]; e++)
            if (c.col[e] < (uint32_t)n) t_col[fill[c.col[e]]++] = (uint32_t)u;
    level_to.prepare(n, t_rowptr, t_col);
    std::vector<int32_t> la, lb;
    for (size_t k = 0; k + 1 < c.merged_cells.size(); k += 2) {
        std::vector<int64_t> sa(c.in_list.begin() + (ptrdiff_t)c.in_ptr[k], c.in_list.begin() + (ptrdiff_t)c.in_ptr[k + 1]);
        std::vector<int64_t> sb(c.in_list.begin() + (ptrdiff_t)c.in_ptr[k + 1], c.in_list.begin() + (ptrdiff_t)c.in_ptr[k + 2]);
        if (sa.empty() || sb.empty()) continue;
        level_to.run(sa, la);
        level_to.run(sb, lb);
        // a source other than the pair itself reaches the cell at level R iff its nearest in-neighbour of that very
        // cell is at level R-1
        for (int64_t s = 0; s < n; s++)
            if (s != c.merged_cells[k] && la[(size_t)s] == radius - 1 && lb[(size_t)s] == radius - 1) {
                total_nodes[(size_t)s] += 1;
                total_depth[(size_t)s] += radius;
                dist[(size_t)s * (size_t)max_levels + (size_t)radius] += 1;
            }
    }
}

template <typename T> inline void copy_from_primary(const std::vector<int32_t> &primary, T *values, size_t width = 1) {
    for (size_t v = 0; v < primary.size(); v++)
        if ((size_t)primary[v] != v)
            std::copy(values + (size_t)primary[v] * width, values + ((size_t)primary[v] + 1) * width, values + v * width);
}

}  // namespace dmx
