// philox.cuh -- counter-based per-chain random streams (Philox4x32-10).
//
// The reference seeds a ChaCha12 StdRng per chain from OS entropy
// (rust/src/lib.rs:206-207), which is irreproducible by construction.  Here a
// chain owns two Philox streams keyed by (seed, global chain id):
//   counter = (block_lo, block_hi, chain_lo, chain_hi | stream << 30), key = seed
//   stream 0 "init": block b holds the two u64 of variables 2b and 2b+1
//   stream 1 "walk": u32 words consumed in the reference's draw order; 64-bit
//                    draws take two words at an even position (an odd cursor
//                    skips one word) so a draw never straddles two blocks and
//                    lane j can fetch the draw of literal j on its own.
// The distributions restate rand 0.8.5 (Standard f64, Bernoulli, UniformInt
// sample_single, UniformFloat) as cited in DESIGN.md.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace slsb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

enum : uint32_t { STREAM_INIT = 0, STREAM_WALK = 1 };

struct ChainKey {
    uint2 key;
    uint32_t chain_lo, chain_hi_stream;
    __device__ __forceinline__ ChainKey(uint64_t seed, uint64_t chain, uint32_t stream)
    {
        key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        chain_lo = (uint32_t)chain;
        chain_hi_stream = ((uint32_t)(chain >> 32) & 0x3fffffffu) | (stream << 30);
    }
    __device__ __forceinline__ uint4 block(uint64_t blk) const
    {
        return philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), chain_lo, chain_hi_stream), key);
    }
    // the u64 at even word position p
    __device__ __forceinline__ uint64_t u64_at(uint64_t p) const
    {
        const uint4 b = block(p >> 2);
        return (p & 2) ? ((uint64_t)b.w << 32 | b.z) : ((uint64_t)b.y << 32 | b.x);
    }
};

// Sequential walk stream; every lane of the warp that owns the chain holds an identical copy.
struct WalkStream {
    ChainKey ck;
    uint64_t pos;
    uint64_t cached;
    uint4 buf;
    __device__ __forceinline__ WalkStream(uint64_t seed, uint64_t chain)
        : ck(seed, chain, STREAM_WALK), pos(0), cached(~0ull), buf(make_uint4(0, 0, 0, 0))
    {
    }
    __device__ __forceinline__ void fill(uint64_t blk)
    {
        if (blk != cached) {
            buf = ck.block(blk);
            cached = blk;
        }
    }
    __device__ __forceinline__ uint32_t next_u32()
    {
        fill(pos >> 2);
        const uint32_t sel = (uint32_t)pos & 3u;
        pos++;
        return sel == 0 ? buf.x : sel == 1 ? buf.y : sel == 2 ? buf.z : buf.w;
    }
    __device__ __forceinline__ uint64_t next_u64()
    {
        pos = (pos + 1) & ~1ull;
        fill(pos >> 2);
        const bool hi = (pos & 2) != 0;
        pos += 2;
        return hi ? ((uint64_t)buf.w << 32 | buf.z) : ((uint64_t)buf.y << 32 | buf.x);
    }
    // rand 0.8.5 Standard<f64>: (next_u64 >> 11) * 2^-53
    __device__ __forceinline__ double next_f64() { return (double)(next_u64() >> 11) * 0x1p-53; }
    // rand 0.8.5 UniformInt<u32>::sample_single(0, range): widening multiply + zone rejection
    __device__ __forceinline__ uint32_t uniform_u32(uint32_t range)
    {
        const uint32_t zone = (range << __clz(range)) - 1u;
        for (;;) {
            const uint32_t v = next_u32();
            const uint32_t lo = v * range;
            if (lo <= zone) return __umulhi(v, range);
        }
    }
    // rand 0.8.5 UniformInt<usize>::sample_single on a 64-bit target
    __device__ __forceinline__ uint64_t uniform_u64(uint64_t range)
    {
        const uint64_t zone = (range << __clzll(range)) - 1ull;
        for (;;) {
            const uint64_t v = next_u64();
            const uint64_t lo = v * range;
            if (lo <= zone) return __umul64hi(v, range);
        }
    }
    // reserve `count` aligned u64 slots; returns the word position of the first one
    __device__ __forceinline__ uint64_t reserve_u64(uint64_t count)
    {
        pos = (pos + 1) & ~1ull;
        const uint64_t p = pos;
        pos += 2 * count;
        return p;
    }
};

// Warp-cooperative view of the same walk stream: the 32 lanes of the warp that owns the chain hold
// 32 consecutive Philox blocks (128 words), lane i the block base/4 + i, so one refill costs each
// lane a single Philox evaluation and a draw is one or two shuffles.  Word positions, alignment
// rules and distributions are exactly those of WalkStream; all methods are warp-collective and
// return warp-uniform values.
struct WarpStream {
    ChainKey ck;
    uint64_t base;  // word position of lane 0's block, multiple of 4
    uint32_t off;   // cursor relative to base
    uint4 blk;
    int lane;
    __device__ __forceinline__ WarpStream(uint64_t seed, uint64_t chain, int lane_)
        : ck(seed, chain, STREAM_WALK), base(0), off(0), lane(lane_)
    {
        blk = ck.block((uint64_t)lane);
    }
    __device__ __forceinline__ uint64_t pos() const { return base + off; }
    // The key passes through an opaque asm so that the compiler cannot hoist the Philox key
    // schedule (18 adds) out of this rarely taken branch into the caller's per-step path.
    __device__ __forceinline__ void refill()
    {
        base += off & ~3u;
        off &= 3u;
        uint2 k = ck.key;
        asm volatile("" : "+r"(k.x), "+r"(k.y));
        const uint64_t b = (base >> 2) + lane;
        blk = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(b >> 32), ck.chain_lo, ck.chain_hi_stream), k);
    }
    __device__ __forceinline__ uint32_t next_u32()
    {
        if (off >= 128u) refill();
        const uint32_t c = off & 3u;
        const uint32_t mine = c == 0 ? blk.x : c == 1 ? blk.y : c == 2 ? blk.z : blk.w;
        const uint32_t v = __shfl_sync(0xffffffffu, mine, off >> 2);
        off++;
        return v;
    }
    // caller guarantees off < 128 (the kernels refill once per step)
    __device__ __forceinline__ uint32_t next_u32_unchecked()
    {
        const uint32_t c = off & 3u;
        const uint32_t mine = c == 0 ? blk.x : c == 1 ? blk.y : c == 2 ? blk.z : blk.w;
        const uint32_t v = __shfl_sync(0xffffffffu, mine, off >> 2);
        off++;
        return v;
    }
    __device__ __forceinline__ uint64_t next_u64()
    {
        off = (off + 1u) & ~1u;
        if (off >= 128u) refill();
        const bool hi_half = (off & 2u) != 0;
        const uint32_t lo = __shfl_sync(0xffffffffu, hi_half ? blk.z : blk.x, off >> 2);
        const uint32_t hi = __shfl_sync(0xffffffffu, hi_half ? blk.w : blk.y, off >> 2);
        off += 2u;
        return (uint64_t)hi << 32 | lo;
    }
    // consume a 64-bit draw whose value is not needed
    __device__ __forceinline__ void skip_u64() { off = ((off + 1u) & ~1u) + 2u; }
    __device__ __forceinline__ double next_f64() { return (double)(next_u64() >> 11) * 0x1p-53; }
    __device__ __forceinline__ uint32_t uniform_u32(uint32_t range)
    {
        const uint32_t zone = (range << __clz(range)) - 1u;
        for (;;) {
            const uint32_t v = next_u32();
            if (v * range <= zone) return __umulhi(v, range);
        }
    }
    __device__ __forceinline__ uint64_t uniform_u64(uint64_t range)
    {
        const uint64_t zone = (range << __clzll(range)) - 1ull;
        for (;;) {
            const uint64_t v = next_u64();
            if (v * range <= zone) return __umul64hi(v, range);
        }
    }
    // reserve `count` aligned u64 slots for per-lane direct generation (ck.u64_at)
    __device__ __forceinline__ uint64_t reserve_u64(uint64_t count)
    {
        off = (off + 1u) & ~1u;
        const uint64_t p = base + off;
        const uint64_t np = p + 2 * count;
        base = np & ~3ull;
        off = (uint32_t)(np & 3ull);
        blk = ck.block((base >> 2) + lane);
        return p;
    }
};

__device__ __forceinline__ double u64_to_f64_53(uint64_t x) { return (double)(x >> 11) * 0x1p-53; }

}  // namespace slsb
