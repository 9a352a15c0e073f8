// State and helpers shared by the translation units of the libmpi shim (mpi_shim.cc: init + collectives over the shm
// mailboxes; mpi_p2p.cc: point-to-point over datagram sockets; mpi_comm.cc: communicators, groups, derived datatypes and the
// collectives of sub-communicators, built on point-to-point).
#pragma once
#include <stddef.h>

#include <string>
#include <vector>

#include "../runtime/rendezvous.h"
#include "mpi.h"

namespace b200mpi_mpi {
extern b200mpi::Rendezvous* g_rv;
extern int g_rank, g_size;
extern bool g_init, g_final;
extern int g_timeout_ms;

// A communicator: an ordered list of world ranks plus a context id that keeps its traffic apart from every other
// communicator this process belongs to. MPI_COMM_WORLD (handle 0, context 0) and MPI_COMM_SELF (handle 1, context 1) always
// exist; MPI_Comm_split / dup / create add entries. `world_like` communicators (all ranks, identity order) use the fast
// shared-memory collectives of mpi_shim.cc, the others the point-to-point based ones of mpi_comm.cc.
struct Comm {
  bool live = false;
  std::vector<int> ranks;   // comm rank -> world rank
  int my = 0;               // this process's rank in the communicator
  int ctx = 0;
  bool world_like = false;
  std::string name;
  int size() const { return (int)ranks.size(); }
};
Comm* comm_of(MPI_Comm c);          // nullptr: not a live communicator (or MPI not initialised)
void comms_reset(bool build);       // MPI_Init (build WORLD and SELF) / MPI_Finalize (drop everything)
size_t derived_type_size(MPI_Datatype t);
bool flatten_type(MPI_Datatype t, size_t count, MPI_Datatype* base, size_t* n);   // contiguous derived type -> base type x n

int fail(const std::string& what);
size_t type_size(MPI_Datatype t);   // bytes of one element, derived (contiguous) types included; 0 = unknown
bool reduce_into(void* acc, const void* x, size_t n, MPI_Datatype t, MPI_Op op);
bool op_supported(MPI_Datatype base_type, MPI_Op op);   // checked by every rank before a reduction moves data
int allgather_bytes(const void* in, void* out, size_t bytes);   // over MPI_COMM_WORLD, shared-memory path
int check(MPI_Comm c);
int p2p_init();        // mpi_p2p.cc: binds this rank's message socket (MPI_Init)
void p2p_shutdown();   // mpi_p2p.cc: closes the message socket (MPI_Finalize)

// mpi_p2p.cc: raw transport. `dest` / `src` are WORLD ranks (src may be MPI_ANY_SOURCE), `ctx` the communicator's context id;
// status->MPI_SOURCE comes back as a world rank (callers translate).
int send_bytes(const void* buf, size_t bytes, int dest, int tag, int ctx);
int recv_bytes(void* buf, size_t cap, int src, int tag, int ctx, MPI_Status* st, bool probe_only, bool blocking, int* flag);

// mpi_comm.cc: collectives of a communicator that is neither world-like nor a singleton (comm-rank addressed, internal tags)
int gen_barrier(Comm* C);
int gen_bcast(Comm* C, void* buf, size_t bytes, int root);
int gen_reduce(Comm* C, const void* send, void* recv, size_t count, MPI_Datatype t, MPI_Op op, int root, bool all);
int gen_allgather(Comm* C, const void* in, void* out, size_t bytes);
int gen_alltoall(Comm* C, const void* in, void* out, size_t bytes);
}  // namespace b200mpi_mpi
