/* fjsph_b200_nccl.h — the slab transport of include/fjsph_b200.h (FjsphCommFn) implemented natively on NCCL:
 * ncclSend / ncclRecv between x-neighbours over NVLink for the ghost and migration exchanges, ncclAllReduce on DEVICE
 * buffers for the step's scalars (residual, npd, time-step maxima: FJSPH_COMM_SUM_DEV / MAX_DEV, one PCIe crossing per
 * scalar).  libfjsph_b200_nccl.so, a small library beside libfjsph_b200.so so that single-GPU hosts need no NCCL.
 *
 * The reference is one shared-memory process (FJSPH.cpp:62); this is what a C++ host -- FJSPH's own main(), see
 * fjsph_b200/csrc/fjsph_run.cpp --ranks N -- uses to run one engine per GPU.  One process per GPU:
 *
 *     char id[FJSPH_NCCL_ID_BYTES];
 *     if (rank == 0) fjsph_nccl_unique_id(id);          // then hand `id` to the other ranks (pipe, file, MPI_Bcast ...)
 *     fjsph_nccl_create(id, rank, world, device, &comm);
 *     fjsph_create(...); fjsph_upload_state(...);       // this rank's particles
 *     fjsph_nccl_attach(comm, engine, x_lo, x_hi);      // = fjsph_set_slab with the NCCL callback
 *     ... fjsph_step ...
 *     fjsph_destroy(engine); fjsph_nccl_destroy(comm);
 */
#ifndef FJSPH_B200_NCCL_H
#define FJSPH_B200_NCCL_H

#include "fjsph_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FJSPH_NCCL_ID_BYTES 256 /* two ncclUniqueIds: one communicator for the collectives, one for the neighbour exchanges */
typedef struct FjsphNcclComm FjsphNcclComm;

int fjsph_nccl_unique_id(char id[FJSPH_NCCL_ID_BYTES]); /* ncclGetUniqueId, on one rank */
int fjsph_nccl_create(const char id[FJSPH_NCCL_ID_BYTES], int32_t rank, int32_t world, int32_t device, FjsphNcclComm** out);
/* fjsph_set_slab(e, rank, world, x_lo, x_hi, <the NCCL callback>, comm) + device-side reductions on */
int fjsph_nccl_attach(FjsphNcclComm* comm, FjsphEngine* e, double x_lo, double x_hi);
/* plain collectives for the host's own bookkeeping (decomposition, gathering frames): doubles, in place, host arrays */
int fjsph_nccl_allreduce_host(FjsphNcclComm* comm, double* v, int64_t n, int32_t op /* FJSPH_COMM_SUM | FJSPH_COMM_MAX */);
int fjsph_nccl_barrier(FjsphNcclComm* comm);
int64_t fjsph_nccl_calls(FjsphNcclComm* comm, int32_t op); /* callback invocations per op since creation */
const char* fjsph_nccl_last_error(void);
int fjsph_nccl_destroy(FjsphNcclComm* comm);

#ifdef __cplusplus
}
#endif
#endif
