// rt_tiles.h — screen-space partition of the image (host side).
//
// The reference renders one thread per pixel on one GPU (Index1D(inLen), RTRenderer.cs:153,205).
// Here every context owns the interleaved screen tiles with (tx + 3*ty) % worldSize == rank
// (worldSize <= 1: all of them).  Owned pixels are enumerated tile by tile and, inside a tile, in
// 8x4 micro-tiles, so that the 32 lanes of a warp trace a compact 8x4 pixel footprint.
#pragma once
#include <stdint.h>

#include <vector>

namespace rtx {

inline bool tile_owned(int tx, int ty, int rank, int worldSize) { return worldSize <= 1 || ((tx + 3 * ty) % worldSize) == rank; }

inline int effective_tile_size(int tileSize) {
    int t = tileSize > 0 ? tileSize : 32;
    t = (t + 7) / 8 * 8;   // multiples of 8 keep micro-tiles inside a tile
    return t;
}

// owned index -> global pixel index (y*width + x); order: owned tiles row-major, micro-tiles row-major, pixels row-major
inline void build_pixel_map(int width, int height, int tileSize, int rank, int worldSize, std::vector<int>& map) {
    map.clear();
    const int T = effective_tile_size(tileSize);
    const int ntx = (width + T - 1) / T, nty = (height + T - 1) / T;
    for (int ty = 0; ty < nty; ty++)
        for (int tx = 0; tx < ntx; tx++) {
            if (!tile_owned(tx, ty, rank, worldSize)) continue;
            const int x0 = tx * T, y0 = ty * T, x1 = x0 + T < width ? x0 + T : width, y1 = y0 + T < height ? y0 + T : height;
            for (int my = y0; my < y1; my += 4)
                for (int mx = x0; mx < x1; mx += 8)
                    for (int y = my; y < my + 4 && y < y1; y++)
                        for (int x = mx; x < mx + 8 && x < x1; x++) map.push_back(y * width + x);
        }
}

inline int64_t count_owned_pixels(int width, int height, int tileSize, int rank, int worldSize) {
    const int T = effective_tile_size(tileSize);
    const int ntx = (width + T - 1) / T, nty = (height + T - 1) / T;
    int64_t n = 0;
    for (int ty = 0; ty < nty; ty++)
        for (int tx = 0; tx < ntx; tx++) {
            if (!tile_owned(tx, ty, rank, worldSize)) continue;
            const int x0 = tx * T, y0 = ty * T, x1 = x0 + T < width ? x0 + T : width, y1 = y0 + T < height ? y0 + T : height;
            n += (int64_t)(x1 - x0) * (y1 - y0);
        }
    return n;
}

}   // namespace rtx
