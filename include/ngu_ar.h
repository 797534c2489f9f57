/*
 * ngu_ar.h — C ABI of the learner-gradient all-reduce over NVLink peer memory (libngu_epi.so).
 *
 * BASELINE.json north_star: "the learner's embedding/RND/R2D2 gradient allreduce runs over NCCL on NVLink".  The
 * reference has no gradient averaging at all; the seams are embedding.py:60-61,
 * lifelong_novelty.py:64-65 and r2d2_learner.py:53-54 (between loss.backward() and clip_grad_norm_).  NCCL is the
 * default transport of ngu_b200.learner_allreduce.GradAllReducer; this is the library's OWN transport for the same
 * call (transport="p2p"), one process per GPU on one NVSwitch box:
 *
 *   every rank's flat gradient bucket lives in a buffer the library allocates and the peers map (CUDA IPC);
 *   one kernel per rank: [ready barrier: flag words stored into every peer] -> rank r averages slice r of the bucket,
 *   reading the slice from all ranks over NVLink (16-byte loads, summed in rank order) and storing the result into all
 *   ranks' buckets (16-byte peer stores) -> [done barrier].  Reduce-scatter and all-gather are ONE pass: each element
 *   crosses NVLink once in and once out, no intermediate buffer, and - every element being summed by exactly one rank
 *   in a fixed order - the result is bit-identical on all ranks.
 *
 * Conventions as in ngu_epi.h (0 / negative NGU_E* codes, `stream` = cudaStream_t as void*).  All ranks must call
 * ngu_ar_allreduce_avg the same number of times, in the same order over their handles.
 */
#ifndef NGU_AR_H
#define NGU_AR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngu_ar ngu_ar;

#define NGU_AR_HANDLE_BYTES 128   /* two CUDA IPC handles: the bucket and the flag words */

/* A bucket of n_floats float32 gradients on `device` (zero-initialised; the size is padded to 16 bytes x world slices). */
int ngu_ar_create(int64_t n_floats, int32_t device, ngu_ar** out);
void ngu_ar_destroy(ngu_ar* h);
const char* ngu_ar_last_error(const ngu_ar* h);
/* The bucket: parameters' .grad tensors are views of it, so backward accumulates straight into the buffer the peers read. */
int ngu_ar_buffer(ngu_ar* h, float** bucket_dev);

/* 1. every rank: ngu_ar_export(h, handle[NGU_AR_HANDLE_BYTES]);  2. all-gather the handles in rank order;
 * 3. every rank: ngu_ar_connect(h, rank, world, handles[world * NGU_AR_HANDLE_BYTES]).  world <= 8, distinct GPUs. */
int ngu_ar_export(ngu_ar* h, void* handle_out);
int ngu_ar_connect(ngu_ar* h, int32_t rank, int32_t world, const void* handles);

/* bucket <- average over ranks, in place, enqueued on `stream` (the gradients must have been written by work that
 * precedes the call on that stream).  Collective. */
int ngu_ar_allreduce_avg(ngu_ar* h, void* stream);
/* Barriers that gave up waiting for a peer (~2 s budget each; must stay 0).  Synchronises the device. */
int64_t ngu_ar_timeouts(ngu_ar* h);

#ifdef __cplusplus
}
#endif
#endif /* NGU_AR_H */
