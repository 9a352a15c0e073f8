// MPI ping-pong between rank 0 and rank 1 (latency for small messages, bandwidth for large ones) plus a token ring over all
// ranks — a generic MPI program of the kind people launch with an MPIJob besides Horovod jobs. Built against the in-tree
// libmpi shim by `make examples`; run as `mpirun -n 2 pingpong` or through examples/mpi-ring/ring.yaml.
#include <mpi.h>

#include <cstdio>
#include <vector>

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank, size;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  char host[MPI_MAX_PROCESSOR_NAME];
  int len;
  MPI_Get_processor_name(host, &len);

  // token ring: rank 0 injects 1, every rank adds its rank + 1 and passes it on
  long long token = 0;
  if (rank == 0) {
    token = 1;
    if (size > 1) {
      MPI_Send(&token, 1, MPI_LONG_LONG, 1, 0, MPI_COMM_WORLD);
      MPI_Recv(&token, 1, MPI_LONG_LONG, size - 1, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    }
    printf("ring of %d ranks closed on %s: token = %lld (expected %lld)\n", size, host, token, 1LL + (long long)size * (size + 1) / 2 - 1);
  } else {
    MPI_Recv(&token, 1, MPI_LONG_LONG, rank - 1, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    token += rank + 1;
    MPI_Send(&token, 1, MPI_LONG_LONG, (rank + 1) % size, 0, MPI_COMM_WORLD);
  }

  if (size >= 2 && rank < 2) {
    for (size_t bytes : {(size_t)8, (size_t)1024, (size_t)65536, (size_t)1 << 20, (size_t)16 << 20}) {
      std::vector<char> buf(bytes, (char)rank);
      const int iters = bytes <= 65536 ? 2000 : (bytes <= (1u << 20) ? 200 : 20);
      for (int phase = 0; phase < 2; phase++) {        // phase 0 warms up
        const double t0 = MPI_Wtime();
        for (int i = 0; i < iters; i++) {
          if (rank == 0) { MPI_Send(buf.data(), (int)bytes, MPI_BYTE, 1, 1, MPI_COMM_WORLD); MPI_Recv(buf.data(), (int)bytes, MPI_BYTE, 1, 1, MPI_COMM_WORLD, MPI_STATUS_IGNORE); }
          else { MPI_Recv(buf.data(), (int)bytes, MPI_BYTE, 0, 1, MPI_COMM_WORLD, MPI_STATUS_IGNORE); MPI_Send(buf.data(), (int)bytes, MPI_BYTE, 0, 1, MPI_COMM_WORLD); }
        }
        const double dt = MPI_Wtime() - t0;
        if (phase == 1 && rank == 0)
          printf("pingpong %9zu bytes: %9.2f us one-way, %8.1f MB/s\n", bytes, dt / iters / 2 * 1e6, bytes / (dt / iters / 2) / 1e6);
      }
    }
  }
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Finalize();
  return 0;
}
