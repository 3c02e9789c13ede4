// deepz_types.cuh -- argument structures of deepz_kernel shared by the host code and the kernel units.
#pragma once
#include <cstddef>
#include "common.cuh"

struct DeepzInput {
    int kind;                 // 0: box grid, 1: explicit zonotopes
    int n;                    // rows of the input sets
    long long N;              // cells of this call
    long long cell_begin;     // box grid: id of the first cell in product order
    long long cyc_blk;        // > 0: block-cyclic ids -- local cell i is cell_begin + (i / cyc_blk) * cyc_stride + i % cyc_blk
    long long cyc_stride;
    int part[64];
    int base[64];
    long long stride[64];     // product of part[0..d-1]
    const double* centers;    // device
    const double* radii;      // device
    const double* C;          // device, n x N
    const long long* col_off; // device, N + 1
    const double* G;          // device, n x col_off[N]
    const int* ncols;         // padded layout (col_off == nullptr): generators of cell i are ncols[i] columns at G + i * pad_cap * n
    int pad_cap;
};

struct DeepzOutput {
    double* out_lo;           // m_out x N or nullptr
    double* out_hi;
    double* hull_lo;          // m_out or nullptr (initialised to +inf / -inf by the host)
    double* hull_hi;
    int keep_cap;
    int* keepN;               // N
    double* keepC;            // m_out x N
    double* keepG;            // m_out x keep_cap x N
    int* liveCount;           // pilot runs: [layer][CLR_MAX_ROWS] cells in which the neuron was left non-zero, or nullptr
};

struct DeepzSched {
    int pass;                 // 0: main (adaptive tiles), 1: retry list, 2: big list (column store in global memory)
    int tile_cells;           // fixed tile size (0 = adaptive)
    int cap;                  // doubles of column store (+ staging) per CTA
    int inplace;              // 1: layer products run in place (every warp owns at most one row block)
    int grow_cols;            // generators every cell is known to gain (sigmoid / tanh layers append one per neuron)
    // tile metadata sized for the launch at hand (dz_geom on the host): the appended-generator masks and the position table
    // live behind DeepzSmem, the column store behind them
    int mask_words;           // 32-bit words of a cell's row mask (ceil(max rows / 32))
    int mask_stride;          // words between the masks of two cells (odd: conflict-free column reads)
    int tile_max;             // cells a tile may hold (<= CLR_TILE_CELLS_MAX)
    int ncol_max;             // columns a tile may hold (<= CLR_NCOL_MAX); the position table has ncol_max + 32 entries
    int store_off;            // byte offset of the column store in dynamic shared memory
    int astage_off;           // byte offset / size of the TMA staging area of wide-K W fragments (0 bytes: none)
    int astage_bytes;
    int fuse_tail;            // 1: the last layer (Id, at most 4 inputs and rows) is applied column by column in the output stage
    long long* retry_list;
    long long* big_list;
    double* workspace;        // pass 2: gridDim.x * cap doubles
};

#define DZ_ITEMS_MAX 384   // pending generators listed per tile (more: the slow walk over the masks)

struct DeepzSmem {
    int nCells, tileT, more;
    int kActNew;                              // being collected by the current relaxation
    int kAct;                                 // rows of the store that can be non-zero: the next product's K (inactive
                                              // neurons at the end of a layer are skipped, see dz_gemm)
    int sMaxP;
    unsigned long long abar;                   // mbarrier of the TMA staging of wide-K W fragments
    long long firstIdx;
    float peakNeed;
    float peakCols;                           // columns per cell at the widest layer of the tile, without the slack
    int colOff[2][CLR_TILE_CELLS_MAX + 1];    // cell offsets of the current and of the next layout
    unsigned short wUsed[2][CLR_TILE_CELLS_MAX];   // columns of a cell that hold data (centre + generators), per layout
    unsigned short uPend[2][CLR_TILE_CELLS_MAX];   // generators of a cell still pending (appended by the last relaxation)
    int nItems;                                    // pending generators of the tile listed in item*
    unsigned short itemCol[DZ_ITEMS_MAX], itemCell[DZ_ITEMS_MAX], itemRow[DZ_ITEMS_MAX];
    long long cellId[CLR_TILE_CELLS_MAX];
    unsigned deadMask[2][CLR_NCOL_MAX / 32];  // generators the reference's remove_zero_columns has dropped
    double hlo[CLR_HULL_MAX], hhi[CLR_HULL_MAX];
    unsigned long long sPin[CLR_MAX_LAYERS], sUnst[CLR_MAX_LAYERS], sNear;
    DevLayer layers[CLR_DEV_MAX_LAYERS];
    // ---- only allocated by the STATS instantiations (see dz_meta_bytes) ----
    unsigned int cellNear[CLR_TILE_CELLS_MAX];
    unsigned short cellPin[CLR_TILE_CELLS_MAX][CLR_MAX_LAYERS];
    unsigned short cellUnst[CLR_TILE_CELLS_MAX][CLR_MAX_LAYERS];
};
__host__ __device__ inline size_t dz_meta_bytes(bool stats) {
    size_t b = stats ? sizeof(DeepzSmem) : offsetof(DeepzSmem, cellNear);
    return (b + 15) & ~(size_t)15;
}

#ifndef DZ_NB
#define DZ_NB 4         // column blocks per group of the layer product (one barrier per group in place)
#endif
#ifndef DZ_MARGIN
#define DZ_MARGIN 1.0625f   // head-room of the adaptive tile size over the growth seen so far (main pass)
#endif
#ifndef DZ_THREADS
#define DZ_THREADS 512   // 16 warps (4 per SM sub-partition: 128 registers per thread)
#endif
#ifndef DZ_CTAS
#define DZ_CTAS 1        // CTAs per SM (experiments: 2 x 256 threads)
#endif
