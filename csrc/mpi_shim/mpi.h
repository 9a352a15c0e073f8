/*
 * mpi.h — the MPI subset provided by libmpi (b200mpi shim).
 *
 * The reference's only native workload (examples/v2beta1/pi/pi.cc:19-52) needs
 * MPI_Init / Comm_rank / Comm_size / Get_processor_name / Reduce / Barrier /
 * Finalize; Horovod-style bootstraps add Bcast / Allreduce / Allgather
 * (SURVEY.md §2.2).  Transport: the job's POSIX-shm rendezvous segment
 * (csrc/runtime/rendezvous.h) — CPU path, no ssh, no network.
 */
#ifndef B200MPI_MPI_H_
#define B200MPI_MPI_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;  /* count_: received bytes */
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Errhandler;
typedef long MPI_Aint;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_COMM_NULL (-1)

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_ERR_ARG 12
#define MPI_ERR_COMM 5
#define MPI_ERR_TYPE 3
#define MPI_ERR_OP 9
#define MPI_ERR_ROOT 7
#define MPI_ERR_GROUP 8

#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL (-1)
#define MPI_PROC_NULL (-2)
#define MPI_UNDEFINED (-32766)
#define MPI_TAG_UB 0x3fffffff
#define MPI_ERR_TRUNCATE 14
#define MPI_ERR_RANK 6
#define MPI_ERR_TAG 4
#define MPI_ERR_REQUEST 19
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_GROUP_EMPTY 0
#define MPI_GROUP_NULL (-1)
#define MPI_DATATYPE_NULL 0
#define MPI_ERRORS_ARE_FATAL 1
#define MPI_ERRORS_RETURN 2
#define MPI_ERRHANDLER_NULL 0
#define MPI_MAX_OBJECT_NAME 128
enum { MPI_IDENT = 0, MPI_CONGRUENT = 1, MPI_SIMILAR = 2, MPI_UNEQUAL = 3 };

enum {
  MPI_CHAR = 1, MPI_SIGNED_CHAR, MPI_UNSIGNED_CHAR, MPI_BYTE, MPI_SHORT, MPI_UNSIGNED_SHORT, MPI_INT, MPI_UNSIGNED,
  MPI_LONG, MPI_UNSIGNED_LONG, MPI_LONG_LONG, MPI_UNSIGNED_LONG_LONG, MPI_FLOAT, MPI_DOUBLE, MPI_INT32_T, MPI_INT64_T,
  MPI_UINT32_T, MPI_UINT64_T, MPI_C_BOOL,
  /* value + index pairs for MPI_MAXLOC / MPI_MINLOC: the C structs { float v; int i; } ... (extent = sizeof the struct) */
  MPI_FLOAT_INT, MPI_DOUBLE_INT, MPI_LONG_INT, MPI_2INT
};
#define MPI_LONG_LONG_INT MPI_LONG_LONG
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_PROD, MPI_LAND, MPI_LOR, MPI_BAND, MPI_BOR, MPI_MAXLOC, MPI_MINLOC };
#define MPI_OP_NULL 0
/* user-defined reductions: inoutvec[i] = invec[i] (op) inoutvec[i]; applied in rank order, so non-commutative functions work */
typedef void(MPI_User_function)(void* invec, void* inoutvec, int* len, MPI_Datatype* datatype);
int MPI_Op_create(MPI_User_function* function, int commute, MPI_Op* op);
int MPI_Op_free(MPI_Op* op);

#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

int MPI_Init(int* argc, char*** argv);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Initialized(int* flag);
int MPI_Finalized(int* flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Get_processor_name(char* name, int* resultlen);
int MPI_Get_version(int* version, int* subversion);
int MPI_Get_library_version(char* version, int* resultlen);
int MPI_Type_size(MPI_Datatype datatype, int* size);
int MPI_Error_string(int errorcode, char* string, int* resultlen);
double MPI_Wtime(void);
double MPI_Wtick(void);
/* predefined attributes of MPI_COMM_WORLD: *(int**)attribute_val points at the value; MPI_APPNUM is the MPMD application
 * context the launcher started this rank in (mpirun prog1 : -np 2 prog2) */
enum { MPI_TAG_UB_KEY = 1, MPI_APPNUM = 2, MPI_UNIVERSE_SIZE = 3, MPI_WTIME_IS_GLOBAL = 4, MPI_HOST = 5, MPI_IO = 6 };
int MPI_Comm_get_attr(MPI_Comm comm, int keyval, void* attribute_val, int* flag);
int MPI_Attr_get(MPI_Comm comm, int keyval, void* attribute_val, int* flag);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, MPI_Comm comm);


/* point-to-point: messages travel as datagrams between per-rank abstract UNIX sockets (mpi_p2p.cc); sends are eager
 * (buffered at the receiver), matching follows the MPI non-overtaking rule per (source, tag, communicator) */
int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Ssend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Bsend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Rsend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Sendrecv_replace(void* buf, int count, MPI_Datatype datatype, int dest, int sendtag, int source, int recvtag, MPI_Comm comm,
                         MPI_Status* status);
int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Wait(MPI_Request* request, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request* requests, MPI_Status* statuses);
int MPI_Test(MPI_Request* request, int* flag, MPI_Status* status);
int MPI_Testall(int count, MPI_Request* requests, int* flag, MPI_Status* statuses);
int MPI_Testany(int count, MPI_Request* requests, int* index, int* flag, MPI_Status* status);
int MPI_Waitany(int count, MPI_Request* requests, int* index, MPI_Status* status);
int MPI_Waitsome(int incount, MPI_Request* requests, int* outcount, int* indices, MPI_Status* statuses);
int MPI_Request_free(MPI_Request* request);
int MPI_Cancel(MPI_Request* request);
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int* flag, MPI_Status* status);
int MPI_Get_count(const MPI_Status* status, MPI_Datatype datatype, int* count);

/* communicators: MPI_Comm_split / dup / create build real sub-communicators (own context id; collectives over the
 * point-to-point layer unless the communicator holds every rank in world order, which keeps the shared-memory paths);
 * MPI_COMM_TYPE_SHARED groups everybody (one box). mpi_comm.cc */
#define MPI_COMM_TYPE_SHARED 1
#define MPI_INFO_NULL 0
typedef int MPI_Info;
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);
int MPI_Comm_split_type(MPI_Comm comm, int split_type, int key, MPI_Info info, MPI_Comm* newcomm);
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* result);
int MPI_Comm_set_name(MPI_Comm comm, const char* name);
int MPI_Comm_get_name(MPI_Comm comm, char* name, int* resultlen);
int MPI_Comm_test_inter(MPI_Comm comm, int* flag);
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler errhandler);   /* errors are always returned to the caller */
int MPI_Comm_get_errhandler(MPI_Comm comm, MPI_Errhandler* errhandler);
int MPI_Errhandler_set(MPI_Comm comm, MPI_Errhandler errhandler);
int MPI_Errhandler_free(MPI_Errhandler* errhandler);
int MPI_Error_class(int errorcode, int* errorclass);
int MPI_Comm_group(MPI_Comm comm, MPI_Group* group);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm* newcomm);
int MPI_Comm_create_group(MPI_Comm comm, MPI_Group group, int tag, MPI_Comm* newcomm);   /* collective over the group's members only */
int MPI_Group_size(MPI_Group group, int* size);
int MPI_Group_rank(MPI_Group group, int* rank);
int MPI_Group_incl(MPI_Group group, int n, const int* ranks, MPI_Group* newgroup);
int MPI_Group_excl(MPI_Group group, int n, const int* ranks, MPI_Group* newgroup);
int MPI_Group_translate_ranks(MPI_Group group1, int n, const int* ranks1, MPI_Group group2, int* ranks2);
int MPI_Group_free(MPI_Group* group);

/* contiguous derived datatypes (usable everywhere a predefined type is, reductions included) */
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* datatype);
int MPI_Type_free(MPI_Datatype* datatype);
int MPI_Type_get_extent(MPI_Datatype datatype, MPI_Aint* lb, MPI_Aint* extent);

/* vector collectives and scans */
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts, const int* displs,
                   MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts, const int* displs,
                MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Reduce_scatter_block(const void* sendbuf, void* recvbuf, int recvcount, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Scan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Exscan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Reduce_scatter(const void* sendbuf, void* recvbuf, const int* recvcounts, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                  const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm comm);

/* nonblocking collectives: the operation completes inside the call (a nonblocking call MAY complete early), the request is
 * born complete - programs written against MPI-3 link and run, without overlap */
int MPI_Ibarrier(MPI_Comm comm, MPI_Request* request);
int MPI_Ibcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm, MPI_Request* request);
int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm, MPI_Request* request);
int MPI_Ireduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm, MPI_Request* request);
int MPI_Iallgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                   MPI_Comm comm, MPI_Request* request);
int MPI_Ialltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                  MPI_Comm comm, MPI_Request* request);
int MPI_Igather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, int root,
                MPI_Comm comm, MPI_Request* request);
int MPI_Iscatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, int root,
                 MPI_Comm comm, MPI_Request* request);

#ifdef __cplusplus
}
#endif
#endif /* B200MPI_MPI_H_ */
