// Tile geometry shared by the sort (sort.cu) and the tile-aware advance (advance_tile.cu).
// Cells (i-1, j-1 of particle_cell, ParticleInCell.jl:28-35) are grouped in 8x8 tiles, tiles in 16x16
// meta-tiles; rows are stored grouped by tile in the order of the tile ORDINAL
//   ord = (((ty>>4)*mtx + (tx>>4)) << 8) | ((ty&15) << 4) | (tx&15),   mtx = ceil(tiles_x/16)
// so that consecutive ordinals are x-neighbours and vertically adjacent tiles stay within ~16 tiles.
#pragma once
#include "pic_device.cuh"

struct TileGeom {
  uint32_t mtx;                // meta-tiles per row
  uint32_t ntiles;             // padded tile count (multiple of 256)
  uint32_t tiles_x, tiles_y;   // real tiles per axis
};

static inline TileGeom tile_geom(const GridDev &g) {
  TileGeom t;
  t.tiles_x = (uint32_t)(g.nx - 1 + 7) / 8;
  t.tiles_y = (uint32_t)(g.ny - 1 + 7) / 8;
  t.mtx = (t.tiles_x + 15) / 16;
  t.ntiles = t.mtx * ((t.tiles_y + 15) / 16) * 256u;
  return t;
}

__host__ __device__ __forceinline__ uint32_t tile_ordinal(uint32_t tx, uint32_t ty, uint32_t mtx) {
  return (((ty >> 4) * mtx + (tx >> 4)) << 8) | ((ty & 15u) << 4) | (tx & 15u);
}
__host__ __device__ __forceinline__ void tile_coords(uint32_t ord, uint32_t mtx, int &tx, int &ty) {
  const uint32_t meta = ord >> 8;
  tx = (int)((meta % mtx) * 16u + (ord & 15u));
  ty = (int)((meta / mtx) * 16u + ((ord >> 4) & 15u));
}

// Direction codes of the incremental re-group (advance_tile.cu): where a row's CURRENT position lies relative
// to the tile it is STORED in.  0..8 = (dy+1)*3 + (dx+1) with dx, dy in {-1,0,1} (4 = same tile),
// 9 = anywhere else (joins the unsorted tail), 10 = discarded row (parked behind the live rows).
constexpr int CODE_STAY = 4, CODE_FAR = 9, CODE_DEAD = 10, NCODE = 16;
