// State and helpers shared by the translation units of the libmpi shim (mpi_shim.cc: init + collectives over the shm
// mailboxes; mpi_p2p.cc: point-to-point over datagram sockets + the collectives built on it).
#pragma once
#include <stddef.h>

#include <string>

#include "../runtime/rendezvous.h"
#include "mpi.h"

namespace b200mpi_mpi {
extern b200mpi::Rendezvous* g_rv;
extern int g_rank, g_size;
extern bool g_init, g_final;
extern int g_timeout_ms;

int fail(const std::string& what);
size_t type_size(MPI_Datatype t);
bool reduce_into(void* acc, const void* x, size_t n, MPI_Datatype t, MPI_Op op);
int allgather_bytes(const void* in, void* out, size_t bytes);
int check(MPI_Comm c);
int p2p_init();        // mpi_p2p.cc: binds this rank's message socket (MPI_Init)
void p2p_shutdown();   // mpi_p2p.cc: closes the message socket (MPI_Finalize)
}  // namespace b200mpi_mpi
