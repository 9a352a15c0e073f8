"""Launch layer: hostfile / discover_hosts.sh generation, per-rank environment,
the mpirun-compatible shim and the native gang spawner."""
