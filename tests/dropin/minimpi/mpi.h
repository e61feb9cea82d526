/* tests/dropin/minimpi/mpi.h -- TEST INFRASTRUCTURE, not part of the product.
 *
 * No MPI is installed in this image, so the reference's executable (src/main.cpp + the three
 * workers + its comm/timer/stats/yaml sources) could not be linked.  This is a from-scratch,
 * single-box, shared-memory implementation of exactly the subset of MPI those sources use
 * (enumerated with `grep -o "MPI_[A-Za-z_0-9]*"` over them), so that the UNMODIFIED reference
 * program can be built twice by tests/dropin/Makefile:
 *
 *   ref_main_cpu   = the reference's program with the reference's own CPU Layer   (runs anywhere)
 *   ref_main_b200  = the same sources linked against this repository's Layer facade (needs a B200)
 *
 * and launched as K processes by tests/dropin/minimpi/minimpirun.py.  Semantics follow the MPI
 * standard for the calls provided (message ordering per (source, context), tag/source wildcards,
 * eager buffered sends, passive-target RMA with atomic fetch-and-op); anything else aborts.
 */
#ifndef MINIMPI_MPI_H
#define MINIMPI_MPI_H
#include <stddef.h>
#include <stdint.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Win;
typedef int MPI_Info;
typedef int MPI_Group;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
  long _bytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_GROUP_NULL (-1)
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_PROC_NULL (-2)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_WTIME_IS_GLOBAL 9

/* basic datatypes: the value is the handle, sizes live in minimpi.cpp */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5
#define MPI_UNSIGNED_LONG_LONG 6
#define MPI_UINT64_T 7
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 9
#define MPI_UNSIGNED 10

#define MPI_SUM 1
#define MPI_NO_OP 2
#define MPI_BOR 3
#define MPI_BAND 4
#define MPI_REPLACE 5
#define MPI_MAX 6
#define MPI_MIN 7

#define MPI_LOCK_EXCLUSIVE 1
#define MPI_LOCK_SHARED 2
#define MPI_MODE_NOCHECK 1024

#define MPI_MAX_PROCESSOR_NAME 256

#ifdef __cplusplus
extern "C" {
#endif
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
double MPI_Wtick(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Get_processor_name(char *name, int *len);

int MPI_Type_create_struct(int n, const int *blocklengths, const MPI_Aint *offsets,
                           const MPI_Datatype *types, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
int MPI_Type_size(MPI_Datatype t, int *size);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype t, int *count);

int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm,
              MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm,
             MPI_Status *status);
int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype st, int dest, int stag, void *rbuf,
                 int rcount, MPI_Datatype rt, int source, int rtag, MPI_Comm comm,
                 MPI_Status *status);
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int *flag, MPI_Status *status);
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status);
int MPI_Wait(MPI_Request *req, MPI_Status *status);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op,
                  MPI_Comm comm);
int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
               MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, const int *rcounts,
                const int *displs, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                  MPI_Datatype rt, MPI_Comm comm);

int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup);
int MPI_Group_free(MPI_Group *group);
int MPI_Comm_create_group(MPI_Comm comm, MPI_Group group, int tag, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);

int MPI_Win_allocate(MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, void *baseptr,
                     MPI_Win *win);
int MPI_Win_free(MPI_Win *win);
int MPI_Win_lock(int lock_type, int rank, int assert_, MPI_Win win);
int MPI_Win_unlock(int rank, MPI_Win win);
int MPI_Put(const void *origin, int ocount, MPI_Datatype ot, int target_rank, MPI_Aint target_disp,
            int tcount, MPI_Datatype tt, MPI_Win win);
int MPI_Get(void *origin, int ocount, MPI_Datatype ot, int target_rank, MPI_Aint target_disp,
            int tcount, MPI_Datatype tt, MPI_Win win);
int MPI_Fetch_and_op(const void *origin, void *result, MPI_Datatype t, int target_rank,
                     MPI_Aint target_disp, MPI_Op op, MPI_Win win);
#ifdef __cplusplus
}
#endif
#endif
