// Monte-Carlo estimate of pi over MPI — the CPU smoke workload of the MPIJob
// examples (same observable behaviour as the reference's
// examples/v2beta1/pi/pi.cc: every rank samples the unit square, rank 0 prints
// "pi is approximately ..."; written independently against our libmpi shim).
#include <mpi.h>

#include <cstdint>
#include <cstdio>
#include <random>

int main(int argc, char** argv) {
  if (MPI_Init(&argc, &argv) != MPI_SUCCESS) return 2;
  int rank = 0, workers = 1, len = 0;
  char host[MPI_MAX_PROCESSOR_NAME];
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &workers);
  MPI_Get_processor_name(host, &len);
  std::printf("Worker %d/%d on %s\n", rank, workers, host);

  const long long samples = argc > 1 ? std::atoll(argv[1]) : 10000000LL;
  std::mt19937_64 gen(0x9E3779B97F4A7C15ull ^ (std::uint64_t)(rank + 1));
  std::uniform_real_distribution<double> unit(0.0, 1.0);
  long long inside = 0;
  for (long long i = 0; i < samples; ++i) {
    const double x = unit(gen), y = unit(gen);
    inside += (x * x + y * y <= 1.0);
  }
  long long total = 0;
  MPI_Reduce(&inside, &total, 1, MPI_LONG_LONG, MPI_SUM, 0, MPI_COMM_WORLD);
  if (rank == 0) std::printf("pi is approximately %.16lf\n", 4.0 * (double)total / ((double)samples * workers));
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Finalize();
  return 0;
}
