// hesaff_b200/csrc/match.cu -- the consumer of the Keypoint records: brute-force nearest-neighbour
// matching of SIFT descriptors, the step that follows detection in the retrieval pipeline the reference is made for
// (README:49-53).  Works on the 164-byte records where they are (device memory; after shard.all_gather_keypoints these are
// the records of every rank), integer arithmetic: squared L2 distance over the 128 descriptor bytes via byte-wise absolute
// differences (__vabsdiffu4) and a 4-way dot product (__dp4a), so results are exact and order independent.
#include "common.cuh"

#define MT_NT 128            // one query per thread
#define MT_DB 64             // database descriptors per shared-memory tile

__global__ void __launch_bounds__(MT_NT) k_match(const hesaff_keypoint *__restrict__ query, uint32_t nq,
                                                 const hesaff_keypoint *__restrict__ db, uint32_t ndb,
                                                 int32_t *__restrict__ best_index, uint32_t *__restrict__ best_d2,
                                                 uint32_t *__restrict__ second_d2)
{
   __shared__ uint32_t s_db[MT_DB][32];
   const uint32_t q = blockIdx.x * MT_NT + threadIdx.x;
   uint32_t qw[32];
   {
      // desc sits at byte 36 of the 164-byte record: 4-byte aligned
      const uint32_t *src = reinterpret_cast<const uint32_t *>(query[min(q, nq - 1)].desc);
#pragma unroll
      for (int i = 0; i < 32; i++) qw[i] = src[i];
   }
   uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;
   int32_t bi = -1;
   for (uint32_t base = 0; base < ndb; base += MT_DB) {
      const uint32_t cnt = min((uint32_t)MT_DB, ndb - base);
      __syncthreads();
      for (uint32_t t = threadIdx.x; t < cnt * 32; t += MT_NT)
         s_db[t >> 5][t & 31] = reinterpret_cast<const uint32_t *>(db[base + (t >> 5)].desc)[t & 31];
      __syncthreads();
      for (uint32_t j = 0; j < cnt; j++) {
         uint32_t d2 = 0;
#pragma unroll
         for (int i = 0; i < 32; i++) {
            const uint32_t ad = __vabsdiffu4(qw[i], s_db[j][i]);
            d2 = __dp4a(ad, ad, d2);
         }
         // ties keep the lower database index (the scan runs in index order)
         if (d2 < b1) { b2 = b1; b1 = d2; bi = (int32_t)(base + j); }
         else if (d2 < b2) b2 = d2;
      }
   }
   if (q < nq) {
      best_index[q] = bi;
      best_d2[q] = b1;
      if (second_d2) second_d2[q] = b2;
   }
}

void ha_launch_match(const hesaff_keypoint *query, uint32_t nq, const hesaff_keypoint *db, uint32_t ndb, int32_t *best_index,
                     uint32_t *best_d2, uint32_t *second_d2, cudaStream_t st)
{
   if (nq == 0) return;
   k_match<<<(nq + MT_NT - 1) / MT_NT, MT_NT, 0, st>>>(query, nq, db, ndb, best_index, best_d2, second_d2);
}
