// rp_device.cuh -- device helpers shared by the CUDA translation units (table probe, posting-block address).
#pragma once
#include "rp_common.h"

namespace rp {

// Cuckoo lookup: both candidate buckets (2 x 2 slots of 16 B) are loaded unconditionally.
// In partitioned mode the owner partition's table is probed (peer memory over NVLink if it is remote).
// one bucket = 2 slots = one 32 B sector: a single 256-bit load (LDG.E.256) instead of two 128-bit ones halves
// the load-pipe work of a probe (32 lanes, 32 different sectors per instruction)
__device__ __forceinline__ void ldg_bucket(const uint4* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ bool table_probe(const DbView& db, uint64_t key, uint64_t& meta) {
  if (db.direct) {  // direct-address table (small nucleotide key spaces)
    meta = __ldg(reinterpret_cast<const unsigned long long*>(db.direct) + key);
    return meta != kEmptyKey;
  }
  const KeyHash m = hash_key(key);
  const int part = db.table_parts > 1 ? (int)owner_of(m, db.table_parts) : 0;
  const uint4* table = db.table[part];
  const int shift = db.bucket_shift[part];
  const uint4* p1 = table + (size_t)bucket1(m, shift) * kBucketSlots;
  const uint4* p2 = table + (size_t)bucket2(m, shift) * kBucketSlots;
  uint4 s0, s1, s2, s3;
  ldg_bucket(p1, s0, s1);
  ldg_bucket(p2, s2, s3);
  const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
  const bool h0 = s0.x == klo && s0.y == khi, h1 = s1.x == klo && s1.y == khi;
  const bool h2 = s2.x == klo && s2.y == khi, h3 = s3.x == klo && s3.y == khi;
  const uint32_t z = h0 ? s0.z : h1 ? s1.z : h2 ? s2.z : s3.z;
  const uint32_t w = h0 ? s0.w : h1 ? s1.w : h2 ? s2.w : s3.w;
  meta = (uint64_t)z | ((uint64_t)w << 32);
  return h0 | h1 | h2 | h3;
}

// posting block a table meta points to (its partition is in the top bits)
__device__ __forceinline__ const uint8_t* block_ptr(const DbView& db, uint64_t meta) {
  return db.blocks[meta >> kMetaPartShift] + ((meta >> 16) & kMetaOffMask) * kBlockAlign;
}


}  // namespace rp
