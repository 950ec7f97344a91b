// tile_geom.h — geometry of the CTA-tiled log-A table (hiC) the persistent engine streams.
//
// Destination columns are split evenly over the G CTAs of the persistent grid (CTA b owns
// [tile_c0(b), tile_c0(b+1))).  A CTA's columns are grouped into rounds of TILE_RW columns (the
// number its consumer warps own at once); within a round the source axis is cut into chunks of
// TILE_CH states, and the table stores, for chunk after chunk, that chunk of every column of the
// round back to back:
//     [CTA b][round][chunk u][column rr of the round][k - u*TILE_CH]
// Every column contributes Kp floats whatever the tiling, so CTA b starts at float tile_c0(b)*Kp,
// a round at +round*TILE_RW*Kp, and one trellis step of a CTA is one linear pass over its part.
// Padding entries (k >= K) hold -inf.
#pragma once

#include <stddef.h>

namespace flashv {

constexpr int TILE_CH = 256;  // source states per chunk (multiple of 128; the last chunk may be shorter)
constexpr int TILE_RW = 28;   // columns per round = consumer warps x columns per warp

__host__ __device__ inline int tile_c0(int K, int G, int b) { return (int)((long long)b * K / G); }

// CTA owning column i.
__host__ __device__ inline int tile_owner(int K, int G, int i)
{
    int b = (int)(((long long)i * G) / K);
    while (b + 1 < G && tile_c0(K, G, b + 1) <= i) ++b;
    while (b > 0 && tile_c0(K, G, b) > i) --b;
    return b;
}

// Offset of (row rr, source k) inside a round that holds ncr columns.
__host__ __device__ inline size_t tile_round_off(int Kp, int ncr, int rr, int k)
{
    const int u = k / TILE_CH;
    const int len = (Kp - u * TILE_CH) < TILE_CH ? (Kp - u * TILE_CH) : TILE_CH;
    return (size_t)u * TILE_CH * ncr + (size_t)rr * len + (size_t)(k - u * TILE_CH);
}

// Offset of (column i, source k) in the whole table.
__host__ __device__ inline size_t tile_off(int K, int Kp, int G, int i, int k)
{
    const int b = tile_owner(K, G, i);
    const int c0 = tile_c0(K, G, b), ncols = tile_c0(K, G, b + 1) - c0;
    const int r = i - c0, rho = r / TILE_RW, rr = r % TILE_RW;
    const int ncr = (ncols - rho * TILE_RW) < TILE_RW ? (ncols - rho * TILE_RW) : TILE_RW;
    return (size_t)c0 * Kp + (size_t)rho * TILE_RW * Kp + tile_round_off(Kp, ncr, rr, k);
}

}  // namespace flashv
