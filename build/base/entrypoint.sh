#!/usr/bin/env bash
# Launcher entrypoint for Hydra-style (Intel MPI / MPICH) jobs — counterpart of the reference's
# build/base/entrypoint.sh:3-36, which blocks the launcher until its own hostname and every host in
# /etc/mpi/hostfile resolve in DNS (exponential back-off 0.1 s x2, <= 10 retries). On a single box
# there is no DNS: a host "resolves" once the node agent has published it in the job's slot map
# (or immediately for CPU-only jobs). Then exec the user command.
set -euo pipefail

resolves() {
  local host="${1%%.*}"
  [[ -z "${B200MPI_SLOTS_FILE:-}" || ! -s "${B200MPI_SLOTS_FILE}" ]] && return 0
  grep -q "\"${host}\"" "${B200MPI_SLOTS_FILE}"
}

wait_for() {
  local host="$1" delay=0.1 tries=0
  until resolves "$host"; do
    tries=$((tries + 1))
    if (( tries > ${B200MPI_ENTRYPOINT_RETRIES:-10} )); then echo "entrypoint: $host never became ready" >&2; return 1; fi
    sleep "$delay"; delay=$(awk "BEGIN{print $delay*2}")
  done
}

if [[ "${K_MPI_JOB_ROLE:-}" == "launcher" ]]; then
  hostfile="${I_MPI_HYDRA_HOST_FILE:-${HYDRA_HOST_FILE:-${OMPI_MCA_orte_default_hostfile:-}}}"
  if [[ -n "$hostfile" && -r "$hostfile" ]]; then
    while read -r line; do
      h="${line%% *}"; h="${h%%:*}"
      [[ -n "$h" ]] && wait_for "$h"
    done < "$hostfile"
  fi
fi
exec "$@"
