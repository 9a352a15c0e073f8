#include "rendezvous.h"

#include <errno.h>
#include <fcntl.h>
#include <linux/futex.h>
#include <sched.h>
#include <signal.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <cstdio>
#include <new>

namespace b200mpi {

static constexpr uint64_t kMagic = 0xB200B200C0117EC7ull;
static constexpr uint32_t kVersion = 3;

uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + ts.tv_nsec;
}

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield");
#endif
}
// futex on a word in the shared segment (no FUTEX_PRIVATE_FLAG: waiters and waker are different processes)
static void futex_wait(std::atomic<uint32_t>* addr, uint32_t expected, long timeout_ns) {
  timespec ts{timeout_ns / 1000000000L, timeout_ns % 1000000000L};
  syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAIT, expected, &ts, nullptr, 0);
}
static void futex_wake_all(std::atomic<uint32_t>* addr) {
  syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAKE, INT32_MAX, nullptr, nullptr, 0);
}

static void nap(int& spins) {
  if (++spins < 200) sched_yield();
  else usleep(spins < 2000 ? 50 : 500);
}

static std::string sanitize(const std::string& s) {
  std::string o;
  for (char ch : s) o.push_back((isalnum((unsigned char)ch) || ch == '-' || ch == '_' || ch == '.') ? ch : '_');
  if (o.size() > 200) o.resize(200);
  return o;
}

Rendezvous::~Rendezvous() { detach(false); }

std::string Rendezvous::sock_name(int rank) const {
  char buf[100];
  snprintf(buf, sizeof(buf), "b200mpi-%016llx-%d", (unsigned long long)(hdr_ ? hdr_->nonce : 0), rank);
  return std::string(buf);
}

int Rendezvous::attach(const std::string& job_id, int rank, int world, int device, int timeout_ms, std::string* err) {
  if (world < 1 || world > kRvMaxRanks || rank < 0 || rank >= world) {
    *err = "rendezvous: invalid rank/world";
    return -EINVAL;
  }
  name_ = "/b200mpi-" + sanitize(job_id);
  rank_ = rank;
  world_ = world;
  const uint64_t t0 = now_ns();
  const size_t size = sizeof(RvHeader);
  if (rank == 0) {
    shm_unlink(name_.c_str());
    int fd = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) { *err = "shm_open(create " + name_ + "): " + strerror(errno); return -errno; }
    if (ftruncate(fd, (off_t)size) != 0) { *err = std::string("ftruncate: ") + strerror(errno); close(fd); return -errno; }
    void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { *err = std::string("mmap: ") + strerror(errno); return -errno; }
    hdr_ = new (p) RvHeader;  // zero pages from ftruncate; atomics are trivially constructed
    hdr_->version = kVersion;
    hdr_->world = world;
    hdr_->creator_pid = getpid();
    hdr_->nonce = now_ns() ^ ((uint64_t)getpid() << 32) ^ 0x9e3779b97f4a7c15ull;
    hdr_->magic.store(kMagic, std::memory_order_release);
  } else {
    int spins = 0;
    for (;;) {
      if ((now_ns() - t0) / 1000000ull > (uint64_t)timeout_ms) {
        *err = "rendezvous: timed out waiting for rank 0 to create " + name_;
        return -ETIMEDOUT;
      }
      int fd = shm_open(name_.c_str(), O_RDWR, 0600);
      if (fd < 0) { nap(spins); continue; }
      struct stat st;
      if (fstat(fd, &st) != 0 || (size_t)st.st_size != size) { close(fd); nap(spins); continue; }
      void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      if (p == MAP_FAILED) { nap(spins); continue; }
      RvHeader* h = reinterpret_cast<RvHeader*>(p);
      bool ok = false;
      for (int k = 0; k < 200; k++) {  // give a live creator 100 ms to finish its header
        if (h->magic.load(std::memory_order_acquire) == kMagic) { ok = true; break; }
        usleep(500);
      }
      // A segment left behind by a crashed run has a dead creator: skip it and
      // wait for the new rank 0 to unlink + recreate.
      if (ok && (h->version != kVersion || h->world != world || kill(h->creator_pid, 0) != 0 ||
                 h->slot[rank].attached.load() != 0))
        ok = false;
      if (!ok) { munmap(p, size); nap(spins); usleep(2000); continue; }
      hdr_ = h;
      break;
    }
  }
  RvSlot& me = hdr_->slot[rank];
  me.pid = getpid();
  me.device = device;
  me.heartbeat_ns.store(now_ns());
  // fd-passing endpoint (abstract namespace: nothing to clean up on the filesystem)
  sock_ = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (sock_ < 0) { *err = std::string("socket: ") + strerror(errno); return -errno; }
  sockaddr_un addr;
  memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  std::string sn = sock_name(rank);
  memcpy(addr.sun_path + 1, sn.data(), sn.size());
  if (bind(sock_, (sockaddr*)&addr, (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + sn.size())) != 0) {
    *err = std::string("bind fd socket: ") + strerror(errno);
    return -errno;
  }
  me.attached.store(1, std::memory_order_release);
  // wait until every rank is attached (and therefore bound)
  int spins = 0;
  for (int r = 0; r < world; r++) {
    while (hdr_->slot[r].attached.load(std::memory_order_acquire) == 0) {
      if ((now_ns() - t0) / 1000000ull > (uint64_t)timeout_ms) {
        *err = "rendezvous: timed out waiting for rank " + std::to_string(r) + " to attach";
        return -ETIMEDOUT;
      }
      nap(spins);
    }
  }
  return barrier(timeout_ms, err);
}

void Rendezvous::detach(bool unlink_segment) {
  if (sock_ >= 0) { close(sock_); sock_ = -1; }
  for (auto& p : stash_) if (p.fd >= 0) close(p.fd);
  stash_.clear();
  if (hdr_) {
    munmap(hdr_, sizeof(RvHeader));
    hdr_ = nullptr;
    if (unlink_segment) shm_unlink(name_.c_str());
  }
}

void Rendezvous::heartbeat() {
  if (hdr_) hdr_->slot[rank_].heartbeat_ns.store(now_ns(), std::memory_order_relaxed);
}

int Rendezvous::barrier(int timeout_ms, std::string* err) {
  if (!hdr_) { *err = "rendezvous: not attached"; return -EINVAL; }
  heartbeat();
  local_sense_ ^= 1u;
  const uint32_t sense = local_sense_;
  if (hdr_->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)world_) {
    hdr_->bar_count.store(0, std::memory_order_relaxed);
    hdr_->bar_sense.store(sense, std::memory_order_release);
    futex_wake_all(&hdr_->bar_sense);
    return 0;
  }
  // Waiters spin briefly (the common case: everyone arrives within microseconds), then sleep in the kernel on the sense word
  // (a futex shared between processes) and are woken by the last arriver; every wake-up or 2 ms they re-check abort / timeout /
  // dead peers. No sched_yield storms when ranks outnumber cores, no busy CPU while a peer is late.
  const uint64_t t0 = now_ns();
  uint64_t last_check = t0;
  int spins = 0;
  while (hdr_->bar_sense.load(std::memory_order_acquire) != sense) {
    if (hdr_->abort_flag.load(std::memory_order_relaxed)) { *err = "rendezvous: job aborted"; return -ECANCELED; }
    if ((++spins & 15) != 0) { cpu_relax(); continue; }
    const uint64_t now = now_ns();
    if (now - t0 < 40000ull) continue;          // spin ~40 us: data-phase skew (a 256 KiB copy) fits in it
    if (now - t0 < 80000ull) { sched_yield(); continue; }
    if (now - last_check > 100000000ull) {    // every 100 ms
      last_check = now;
      if ((now - t0) / 1000000ull > (uint64_t)timeout_ms) { *err = "rendezvous: host barrier timed out"; return -ETIMEDOUT; }
      for (int r = 0; r < world_; r++) {
        if (r != rank_ && kill(hdr_->slot[r].pid, 0) != 0 && errno == ESRCH) {
          *err = "rendezvous: rank " + std::to_string(r) + " (pid " + std::to_string(hdr_->slot[r].pid) + ") died";
          hdr_->abort_flag.store(1);
          futex_wake_all(&hdr_->bar_sense);
          return -EPIPE;
        }
      }
      heartbeat();
    }
    futex_wait(&hdr_->bar_sense, sense ^ 1u, 2000000);   // returns at once if the word already changed
  }
  return 0;
}

int Rendezvous::allgather(const void* in, void* out, size_t bytes, int timeout_ms, std::string* err) {
  if (bytes > kRvMailbox) { *err = "rendezvous: allgather payload too large"; return -EINVAL; }
  memcpy(hdr_->slot[rank_].mailbox, in, bytes);
  int rc = barrier(timeout_ms, err);
  if (rc) return rc;
  for (int r = 0; r < world_; r++) memcpy((char*)out + (size_t)r * bytes, hdr_->slot[r].mailbox, bytes);
  return barrier(timeout_ms, err);
}

int Rendezvous::bcast(void* buf, size_t bytes, int root, int timeout_ms, std::string* err) {
  size_t done = 0;
  while (done < bytes || bytes == 0) {
    const size_t n = bytes - done < kRvMailbox ? bytes - done : kRvMailbox;
    if (rank_ == root) memcpy(hdr_->slot[root].mailbox, (char*)buf + done, n);
    int rc = barrier(timeout_ms, err);
    if (rc) return rc;
    if (rank_ != root) memcpy((char*)buf + done, hdr_->slot[root].mailbox, n);
    rc = barrier(timeout_ms, err);
    if (rc) return rc;
    done += n;
    if (bytes == 0) break;
  }
  return 0;
}

unsigned char* Rendezvous::open_boxes(size_t box, int timeout_ms, size_t* total_bytes) {
  if (!hdr_ || box == 0) return nullptr;
  const size_t bytes = (size_t)world_ * 2 * box;
  char name[96];
  snprintf(name, sizeof(name), "/b200mpi-boxes-%016llx", (unsigned long long)hdr_->nonce);
  std::string err;
  int fd = -1;
  if (rank_ == 0) {
    shm_unlink(name);
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    // fallocate: fail here (and fall back) rather than SIGBUS later when /dev/shm is too small
    if (fd >= 0 && (ftruncate(fd, (off_t)bytes) != 0 || posix_fallocate(fd, 0, (off_t)bytes) != 0)) { close(fd); fd = -1; shm_unlink(name); }
  }
  std::vector<unsigned char> all((size_t)world_);
  unsigned char ok = fd >= 0;
  if (allgather(&ok, all.data(), 1, timeout_ms, &err)) { if (fd >= 0) { close(fd); shm_unlink(name); } return nullptr; }
  if (!all[0]) return nullptr;                       // rank 0 could not create it
  if (rank_ != 0) fd = shm_open(name, O_RDWR, 0600);
  void* m = fd >= 0 ? mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0) : MAP_FAILED;
  if (fd >= 0) close(fd);
  ok = m != MAP_FAILED;
  const int rc = allgather(&ok, all.data(), 1, timeout_ms, &err);   // has everyone mapped it?
  if (rank_ == 0) shm_unlink(name);
  bool every = rc == 0;
  for (unsigned char v : all) every = every && v;
  if (!every) { if (m != MAP_FAILED) munmap(m, bytes); return nullptr; }
  if (total_bytes) *total_bytes = bytes;
  return static_cast<unsigned char*>(m);
}

void Rendezvous::close_boxes(unsigned char* base, size_t total_bytes) {
  if (base) munmap(base, total_bytes);
}

struct FdMsg { uint32_t src; uint32_t tag; };

int Rendezvous::send_fd(int dst, uint32_t tag, int fd, std::string* err) {
  sockaddr_un addr;
  memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  std::string sn = sock_name(dst);
  memcpy(addr.sun_path + 1, sn.data(), sn.size());
  FdMsg m{(uint32_t)rank_, tag};
  iovec iov{&m, sizeof(m)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_name = &addr;
  msg.msg_namelen = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + sn.size());
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  for (int attempt = 0; attempt < 2000; attempt++) {
    if (sendmsg(sock_, &msg, 0) >= 0) return 0;
    if (errno != EAGAIN && errno != ENOBUFS && errno != ECONNREFUSED && errno != EINTR) break;
    usleep(1000);
  }
  *err = std::string("send_fd: ") + strerror(errno);
  return -errno;
}

int Rendezvous::recv_fd(int src, uint32_t tag, int timeout_ms, int* fd, std::string* err) {
  for (size_t i = 0; i < stash_.size(); i++) {
    if (stash_[i].src == src && stash_[i].tag == tag) {
      *fd = stash_[i].fd;
      stash_.erase(stash_.begin() + i);
      return 0;
    }
  }
  const uint64_t t0 = now_ns();
  timeval tv{0, 100000};
  setsockopt(sock_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  for (;;) {
    FdMsg m{};
    iovec iov{&m, sizeof(m)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    ssize_t n = recvmsg(sock_, &msg, MSG_CMSG_CLOEXEC);
    if (n < 0) {
      if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) {
        if (hdr_->abort_flag.load()) { *err = "recv_fd: job aborted"; return -ECANCELED; }
        if ((now_ns() - t0) / 1000000ull > (uint64_t)timeout_ms) { *err = "recv_fd: timed out"; return -ETIMEDOUT; }
        continue;
      }
      *err = std::string("recv_fd: ") + strerror(errno);
      return -errno;
    }
    int got = -1;
    for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c))
      if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) memcpy(&got, CMSG_DATA(c), sizeof(int));
    if (n != (ssize_t)sizeof(m) || got < 0) continue;
    if ((int)m.src == src && m.tag == tag) { *fd = got; return 0; }
    stash_.push_back(Pending{(int)m.src, m.tag, got});
  }
}

}  // namespace b200mpi
