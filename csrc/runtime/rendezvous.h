// Shared-memory rendezvous + fd passing for ranks of one job on one box.
//
// Replaces what the reference gets from ssh + headless-Service DNS + the
// ncclUniqueId broadcast over MPI: ranks meet in a POSIX
// shm segment named after the job, run host barriers / small allgathers there,
// publish heartbeats for the daemon's failure detector and
// hand each other CUDA VMM file descriptors over abstract UNIX datagram
// sockets (SCM_RIGHTS).
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200mpi {

constexpr int kRvMaxRanks = 64;       // CPU-only jobs (libmpi shim) may exceed 8 ranks
constexpr size_t kRvScratch = 256;    // per-rank payload of host_allgather
constexpr size_t kRvMailbox = 1 << 16;  // per-rank bulk mailbox (CPU collectives)

struct RvSlot {
  std::atomic<uint32_t> attached;
  int32_t pid;
  int32_t device;
  uint32_t caps;
  std::atomic<uint64_t> heartbeat_ns;
  std::atomic<uint64_t> mail_seq;   // sequence number of the mailbox content
  std::atomic<uint64_t> mail_ack;   // consumers done with mail_seq
  uint64_t mail_bytes;
  alignas(64) unsigned char scratch[kRvScratch];
  alignas(64) unsigned char mailbox[kRvMailbox];
};

struct RvHeader {
  std::atomic<uint64_t> magic;
  uint32_t version;
  int32_t world;
  int32_t creator_pid;
  uint64_t nonce;
  std::atomic<uint32_t> bar_count;
  std::atomic<uint32_t> bar_sense;
  std::atomic<uint32_t> abort_flag;
  std::atomic<uint32_t> generation;  // bumped by the daemon on elastic rescale
  alignas(64) RvSlot slot[kRvMaxRanks];
};

class Rendezvous {
 public:
  Rendezvous() = default;
  ~Rendezvous();
  // Attach (rank 0 creates). Returns 0 or negative errno-style code; err filled.
  int attach(const std::string& job_id, int rank, int world, int device, int timeout_ms, std::string* err);
  void detach(bool unlink_segment);
  int barrier(int timeout_ms, std::string* err);
  int allgather(const void* in, void* out, size_t bytes, int timeout_ms, std::string* err);
  // bulk broadcast/gather helpers used by the CPU libmpi shim
  int bcast(void* buf, size_t bytes, int root, int timeout_ms, std::string* err);
  // Bulk-data boxes for host collectives: per rank a {data, result} pair of `box` bytes in a segment of its own, created by
  // rank 0, mapped by every rank, then unlinked at once (the mappings keep it alive; nothing is left behind after a crash).
  // Collective call. Returns the base address (rank r's data box = base + r * 2 * box, its result box follows) or nullptr when
  // any rank could not create / map it (e.g. /dev/shm too small): callers then keep using the 64 KiB mailboxes.
  unsigned char* open_boxes(size_t box, int timeout_ms, size_t* total_bytes);
  static void close_boxes(unsigned char* base, size_t total_bytes);
  void heartbeat();
  void set_abort() { if (hdr_) hdr_->abort_flag.store(1); }
  bool aborted() const { return hdr_ && hdr_->abort_flag.load() != 0; }
  // fd passing: send `fd` with tag to `dst`; receive the fd sent by `src` with `tag`.
  int send_fd(int dst, uint32_t tag, int fd, std::string* err);
  int recv_fd(int src, uint32_t tag, int timeout_ms, int* fd, std::string* err);
  int rank() const { return rank_; }
  int world() const { return world_; }
  RvHeader* header() { return hdr_; }
  const std::string& name() const { return name_; }

 private:
  std::string sock_name(int rank) const;
  RvHeader* hdr_ = nullptr;
  std::string name_;
  int rank_ = -1, world_ = 0;
  int sock_ = -1;
  uint32_t local_sense_ = 0;
  struct Pending { int src; uint32_t tag; int fd; };
  std::vector<Pending> stash_;
};

uint64_t now_ns();

}  // namespace b200mpi
