// b200mpi-mpirun: native single-box replacement for mpirun / orterun / Hydra
// mpiexec, the launcher every reference example invokes
// (examples/v2beta1/pi/pi.yaml:21-26,
//  examples/v2beta1/tensorflow-benchmarks/tensorflow-benchmarks.yaml:17-42).
//
// The reference reaches workers over ssh and lets orted fork the ranks
// (SURVEY.md §3.3). On one NVSwitch box there is nothing to ssh into: this
// binary reads the same hostfile the controller generated, places ranks "by
// slot", gang-spawns them as local process groups with the per-rank
// environment of every dialect workloads read (OMPI_COMM_WORLD_*, PMI_*,
// RANK/WORLD_SIZE/LOCAL_RANK, HOROVOD_*), pins GPUs from the node agent's slot
// map, multiplexes stdio, forwards signals and propagates the first failure —
// the contract mpirun gives the launcher pod.
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <sched.h>
#include <signal.h>
#include <string.h>
#include <sys/syscall.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <dirent.h>
#include <string>
#include <sys/stat.h>
#include <vector>

extern char** environ;

struct Host { std::string name; int slots; };
struct Rank {
  int rank = 0, local_rank = 0, node = 0;
  std::string host;
  pid_t pid = -1;
  int pidfd = -1, out = -1, err = -1;
  bool exited = false;
  bool retired = false;   // elastic scale-down: this rank is expected to leave; its exit is not a failure
  int status = 0;
  std::string obuf, ebuf;
};

static volatile sig_atomic_t g_signal = 0;
static void on_signal(int s) { g_signal = s; }

static void die(const std::string& m, int code = 1) {
  fprintf(stderr, "mpirun (b200mpi): %s\n", m.c_str());
  exit(code);
}

static std::string short_host(const std::string& fqdn) {
  size_t d = fqdn.find('.');
  return d == std::string::npos ? fqdn : fqdn.substr(0, d);
}


// ---- processor binding (Open MPI's --bind-to none|numa|socket|core|hwthread, --cpus-per-proc, --cpu-set, --report-bindings).
// The default is none (every reference command line says `-bind-to none`). What "numa" means on this box: the rank runs on the
// CPUs of the NUMA node its GPU hangs off (sysfs: /sys/bus/pci/devices/<bus id>/numa_node), so its pinned staging buffers are
// first-touched in memory local to the GPU's PCIe root and the H2D copy does not cross the socket interconnect; without GPUs the
// ranks are spread over the NUMA nodes in blocks. All paths can be re-rooted with B200MPI_SYSFS_ROOT (tests).
enum BindPolicy { BIND_NONE = 0, BIND_NUMA, BIND_CORE };
static std::string sysroot() { const char* r = getenv("B200MPI_SYSFS_ROOT"); return r ? r : ""; }
static std::string read_first_line(const std::string& path) {
  std::ifstream f(path);
  std::string line;
  if (f) std::getline(f, line);
  return line;
}
static std::vector<int> parse_cpulist(const std::string& text) {   // "0-3,8,10-11"
  std::vector<int> out;
  std::istringstream is(text);
  std::string tok;
  while (std::getline(is, tok, ',')) {
    if (tok.empty()) continue;
    int a = 0, b = 0;
    const int n = sscanf(tok.c_str(), "%d-%d", &a, &b);
    if (n == 1) b = a;
    if (n >= 1 && a >= 0 && b >= a && b < CPU_SETSIZE) for (int c = a; c <= b; c++) out.push_back(c);
  }
  return out;
}
static std::string format_cpulist(const std::vector<int>& cpus) {
  std::string out;
  for (size_t k = 0; k < cpus.size();) {
    size_t e = k;
    while (e + 1 < cpus.size() && cpus[e + 1] == cpus[e] + 1) e++;
    out += (out.empty() ? "" : ",") + std::to_string(cpus[k]) + (e > k ? "-" + std::to_string(cpus[e]) : "");
    k = e + 1;
  }
  return out;
}
static std::vector<int> allowed_cpus(const std::string& cpu_set) {
  cpu_set_t m;
  CPU_ZERO(&m);
  std::vector<int> out;
  if (sched_getaffinity(0, sizeof(m), &m) == 0) for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &m)) out.push_back(c);
  if (!cpu_set.empty()) {
    const std::vector<int> want = parse_cpulist(cpu_set);
    std::vector<int> both;
    for (int c : out) if (std::find(want.begin(), want.end(), c) != want.end()) both.push_back(c);
    out = both;
  }
  return out;
}
// NUMA nodes with their CPUs (restricted to `allowed`); nodes without allowed CPUs are dropped
static std::vector<std::pair<int, std::vector<int>>> numa_nodes(const std::vector<int>& allowed) {
  std::vector<std::pair<int, std::vector<int>>> nodes;
  const std::string dir = sysroot() + "/sys/devices/system/node";
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d)) {
      int id = -1;
      if (sscanf(e->d_name, "node%d", &id) != 1 || id < 0) continue;
      std::vector<int> cpus;
      for (int c : parse_cpulist(read_first_line(dir + "/" + e->d_name + "/cpulist")))
        if (std::find(allowed.begin(), allowed.end(), c) != allowed.end()) cpus.push_back(c);
      if (!cpus.empty()) nodes.push_back({id, cpus});
    }
    closedir(d);
  }
  std::sort(nodes.begin(), nodes.end());
  if (nodes.empty() && !allowed.empty()) nodes.push_back({0, allowed});
  return nodes;
}
// NUMA node of GPU `index` (nvidia-smi / PCI bus order): the index-th entry of /proc/driver/nvidia/gpus sorted by bus id
static int gpu_numa_node(int index) {
  std::vector<std::string> bus;
  const std::string dir = sysroot() + "/proc/driver/nvidia/gpus";
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d)) if (e->d_name[0] != '.') bus.push_back(e->d_name);
    closedir(d);
  }
  std::sort(bus.begin(), bus.end());
  if (index < 0 || index >= (int)bus.size()) return -1;
  std::string id = bus[index];
  std::transform(id.begin(), id.end(), id.begin(), ::tolower);
  const std::string v = read_first_line(sysroot() + "/sys/bus/pci/devices/" + id + "/numa_node");
  return v.empty() ? -1 : atoi(v.c_str());
}
// CPUs of local rank `lrank` of `lsize` under `policy` (gpus[r] = GPU index of local rank r, or absent); empty = leave it unbound
static std::vector<int> binding_for(BindPolicy policy, int lrank, int lsize, const std::vector<int>& gpus, int cpus_per_proc,
                                    const std::string& cpu_set, int* numa_out) {
  *numa_out = -1;
  if (policy == BIND_NONE) return {};
  const auto nodes = numa_nodes(allowed_cpus(cpu_set));
  if (nodes.empty()) return {};
  auto node_of = [&](int r) -> size_t {
    const int gnode = r < (int)gpus.size() && gpus[r] >= 0 ? gpu_numa_node(gpus[r]) : -1;
    for (size_t k = 0; k < nodes.size() && gnode >= 0; k++) if (nodes[k].first == gnode) return k;
    return std::min((size_t)r * nodes.size() / (size_t)std::max(1, lsize), nodes.size() - 1);   // blocks of consecutive ranks per node
  };
  const size_t pick = node_of(lrank);
  *numa_out = nodes[pick].first;
  const std::vector<int>& cpus = nodes[pick].second;
  if (policy == BIND_NUMA) return cpus;
  // core: `cpus_per_proc` consecutive CPUs of the node; the ranks that share a node take consecutive chunks, wrapping when the
  // node is oversubscribed
  int before = 0;
  for (int r = 0; r < lrank; r++) if (node_of(r) == pick) before++;
  const int per = std::max(1, cpus_per_proc), chunks = std::max(1, (int)cpus.size() / per);
  const int chunk = before % chunks;
  std::vector<int> out;
  for (int k = 0; k < per && chunk * per + k < (int)cpus.size(); k++) out.push_back(cpus[chunk * per + k]);
  return out;
}

static std::vector<Host> read_hostfile(const std::string& path, int default_slots) {
  std::vector<Host> hosts;
  std::ifstream f(path);
  if (!f) die("cannot open hostfile " + path);
  std::string line;
  while (std::getline(f, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    std::istringstream is(line);
    std::string name, tok;
    if (!(is >> name)) continue;
    int slots = default_slots;
    size_t colon = name.find(':');  // Hydra "host:n"
    if (colon != std::string::npos) {
      slots = atoi(name.c_str() + colon + 1);
      name = name.substr(0, colon);
    }
    while (is >> tok) {  // Open MPI "host slots=n max_slots=m"
      if (tok.rfind("slots=", 0) == 0) slots = atoi(tok.c_str() + 6);
    }
    hosts.push_back({name, slots > 0 ? slots : 1});
  }
  return hosts;
}

// minimal parser for {"hosts": {"name": [0,1], ...}}
static std::map<std::string, std::vector<int>> read_slots_file(const char* path) {
  std::map<std::string, std::vector<int>> out;
  if (!path || !*path) return out;
  std::ifstream f(path);
  if (!f) return out;
  std::stringstream ss;
  ss << f.rdbuf();
  std::string s = ss.str();
  size_t p = s.find("\"hosts\"");
  if (p == std::string::npos) return out;
  p = s.find('{', p);
  if (p == std::string::npos) return out;
  ++p;
  while (p < s.size()) {
    size_t q1 = s.find('"', p);
    if (q1 == std::string::npos) break;
    size_t close = s.find('}', p);
    if (close != std::string::npos && close < q1) break;
    size_t q2 = s.find('"', q1 + 1);
    std::string name = s.substr(q1 + 1, q2 - q1 - 1);
    size_t b1 = s.find('[', q2), b2 = s.find(']', b1);
    if (b1 == std::string::npos || b2 == std::string::npos) break;
    std::vector<int> v;
    std::string nums = s.substr(b1 + 1, b2 - b1 - 1);
    std::istringstream is(nums);
    std::string t;
    while (std::getline(is, t, ',')) {
      bool digit = false;
      for (char c : t) digit = digit || isdigit((unsigned char)c);
      if (digit) v.push_back(atoi(t.c_str()));
    }
    out[name] = v;
    p = b2 + 1;
  }
  return out;
}

// -output-filename DIR (Open MPI): besides the multiplexed console, every rank's streams go to DIR/1/rank.<N>/{stdout,stderr};
// -timestamp-output: every line is prefixed with the wall-clock time it was forwarded.
static std::string g_outdir;
static bool g_timestamp = false;
static bool g_merge_stderr = false;   // -merge-stderr-to-stdout
static int g_stdin_rank = 0;          // -stdin RANK|all|none: which rank keeps the launcher's stdin (-1 none, -2 all)
static FILE* rank_file(int rank, const char* chan) {
  static std::map<std::pair<int, std::string>, FILE*> files;
  auto key = std::make_pair(rank, std::string(chan));
  auto it = files.find(key);
  if (it != files.end()) return it->second;
  std::string dir = g_outdir + "/1";
  mkdir(g_outdir.c_str(), 0755);
  mkdir(dir.c_str(), 0755);
  dir += "/rank." + std::to_string(rank);
  mkdir(dir.c_str(), 0755);
  FILE* f = fopen((dir + "/" + chan).c_str(), "w");
  files[key] = f;
  return f;
}
static void emit_line(FILE* to, bool tag, int rank, const char* chan, const char* data, size_t n, bool add_newline) {
  if (g_timestamp) {
    char ts[40];
    time_t now = time(nullptr);
    struct tm tmv;
    localtime_r(&now, &tmv);
    strftime(ts, sizeof(ts), "%a %b %d %H:%M:%S %Y", &tmv);
    fprintf(to, "%s<%s>:", ts, chan);
  }
  if (tag) fprintf(to, "[1,%d]<%s>:", rank, chan);
  fwrite(data, 1, n, to);
  if (add_newline) fputc('\n', to);
  if (!g_outdir.empty()) {
    if (FILE* f = rank_file(rank, chan)) { fwrite(data, 1, n, f); if (add_newline) fputc('\n', f); fflush(f); }
  }
}
static void flush_lines(std::string& buf, FILE* to, bool tag, int rank, const char* chan, bool final) {
  size_t pos;
  while ((pos = buf.find('\n')) != std::string::npos) {
    emit_line(to, tag, rank, chan, buf.data(), pos + 1, false);
    buf.erase(0, pos + 1);
  }
  if (final && !buf.empty()) {
    emit_line(to, tag, rank, chan, buf.data(), buf.size(), true);
    buf.clear();
  }
  fflush(to);
}

static bool drain(int& fd, std::string& buf) {
  char tmp[65536];
  for (;;) {
    ssize_t n = read(fd, tmp, sizeof(tmp));
    if (n > 0) { buf.append(tmp, n); continue; }
    if (n == 0) { close(fd); fd = -1; return false; }
    if (errno == EAGAIN || errno == EWOULDBLOCK) return true;
    if (errno == EINTR) continue;
    close(fd); fd = -1; return false;
  }
}

int main(int argc, char** argv) {
  int np = -1, ppn = -1, timeout_s = 0, cpus_per_proc = 1;
  bool tag_output = false, oversubscribe = false, report_bindings = false, bind_explicit = false;
  BindPolicy bind_policy = BIND_NONE;
  bool pe_given = false;     // -map-by ...:PE=n implies core binding unless a -bind-to says otherwise (Open MPI)
  std::string cpu_set;
  auto set_bind = [&](std::string v) {
    const size_t colon = v.find(':');           // "core:overload-allowed"
    if (colon != std::string::npos) v = v.substr(0, colon);
    if (v == "none" || v == "board") bind_policy = BIND_NONE;
    else if (v == "numa" || v == "socket" || v == "package" || v == "l3cache" || v == "l2cache" || v == "l1cache") bind_policy = BIND_NUMA;
    else if (v == "core" || v == "hwthread") bind_policy = BIND_CORE;
    else die("unknown -bind-to policy " + v + " (none, numa, socket, core, hwthread)");
  };
  if (const char* b = getenv("B200MPI_BIND_TO")) if (*b) { set_bind(b); bind_explicit = true; }
  std::string hostfile, hostlist, wdir;
  std::vector<std::pair<std::string, std::string>> xenv;  // -x / -genv / -env
  std::vector<std::string> passthrough_unset;
  int i = 1;
  auto need = [&](int k) { if (i + k >= argc) die(std::string("option ") + argv[i] + " needs an argument"); };
  for (; i < argc; i++) {
    std::string a = argv[i];
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') a = a.substr(1);  // --np == -np
    if (a == "-n" || a == "-np" || a == "-c") { need(1); np = atoi(argv[++i]); }
    else if (a == "-allow-run-as-root" || a == "-oversubscribe" || a == "-q" || a == "-quiet" || a == "-nooversubscribe" ||
             a == "-display-map" || a == "-display-allocation" || a == "-report-bindings" || a == "-keep-fqdn-hostnames" ||
             a == "-enable-recovery" || a == "-disable-recovery" || a == "-l") {
      if (a == "-oversubscribe") oversubscribe = true;
      if (a == "-report-bindings") report_bindings = true;
      if (a == "-l") tag_output = true;
    }
    else if (a == "-use-hwthread-cpus" || a == "-verbose" || a == "-v" || a == "-d" || a == "-debug-devel" || a == "-debug-daemons" ||
             a == "-leave-session-attached" || a == "-continuous" || a == "-no-daemonize" || a == "-do-not-launch" || a == "-novm" ||
             a == "-bynode" || a == "-byslot" || a == "-bycore" || a == "-bysocket" || a == "-nolocal" || a == "-tag-output-full") {}
    else if (a == "-pernode") ppn = 1;
    else if (a == "-merge-stderr-to-stdout") g_merge_stderr = true;
    else if (a == "-stdin") {
      need(1);
      const std::string v = argv[++i];
      g_stdin_rank = v == "none" ? -1 : v == "all" ? -2 : atoi(v.c_str());
    }
    else if (a == "-output-directory") { need(1); g_outdir = argv[++i]; }
    else if (a == "-tune" || a == "-am" || a == "-report-uri" || a == "-xterm" || a == "-max-restarts" || a == "-max-vm-size" ||
             a == "-slot-list" || a == "-rankfile" || a == "-rf" || a == "-ompi-server" || a == "-path") { need(1); ++i; }
    else if (a == "-tag-output" || a == "-prepend-rank") tag_output = true;
    else if (a == "-timestamp-output") g_timestamp = true;
    else if (a == "-output-filename" || a == "-outfile-pattern") { need(1); g_outdir = argv[++i]; }
    else if (a == "-bind-to") { need(1); set_bind(argv[++i]); bind_explicit = true; }
    else if (a == "-bind-to-core") { bind_policy = BIND_CORE; bind_explicit = true; }
    else if (a == "-bind-to-socket") { bind_policy = BIND_NUMA; bind_explicit = true; }
    else if (a == "-bind-to-none") { bind_policy = BIND_NONE; bind_explicit = true; }
    else if (a == "-cpus-per-proc" || a == "-cpus-per-rank") { need(1); cpus_per_proc = std::max(1, atoi(argv[++i])); }
    else if (a == "-cpu-set") { need(1); cpu_set = argv[++i]; }
    else if (a == "-map-by") {
      need(1);
      const std::string v = argv[++i];          // "slot", "numa:PE=4", "ppr:1:node" ... only the PE= modifier matters on one box
      const size_t pe = v.find("PE=");
      if (pe != std::string::npos) { cpus_per_proc = std::max(1, atoi(v.c_str() + pe + 3)); pe_given = true; }
    }
    else if (a == "-rank-by" || a == "-prefix" || a == "-launcher" || a == "-launcher-exec" ||
             a == "-bootstrap" || a == "-bootstrap-exec" || a == "-iface" || a == "-report-pid") { need(1); ++i; }
    else if (a == "-mca" || a == "-gmca" || a == "-omca" || a == "-pmixmca" || a == "-prtemca") {
      need(2);
      std::string k = argv[i + 1], v = argv[i + 2];
      xenv.push_back({"OMPI_MCA_" + k, v});  // what orted exports for -mca
      i += 2;
    }
    else if (a == "-x") {
      need(1);
      std::string kv = argv[++i];
      size_t eq = kv.find('=');
      if (eq != std::string::npos) xenv.push_back({kv.substr(0, eq), kv.substr(eq + 1)});
      else if (const char* v = getenv(kv.c_str())) xenv.push_back({kv, v});
    }
    else if (a == "-genv" || a == "-env") { need(2); xenv.push_back({argv[i + 1], argv[i + 2]}); i += 2; }
    else if (a == "-genvall" || a == "-envall") {}
    else if (a == "-hostfile" || a == "-machinefile" || a == "-f" || a == "-default-hostfile") { need(1); hostfile = argv[++i]; }
    else if (a == "-host" || a == "-H" || a == "-hosts") { need(1); hostlist = argv[++i]; }
    else if (a == "-ppn" || a == "-perhost" || a == "-npernode" || a == "-N") { need(1); ppn = atoi(argv[++i]); }
    else if (a == "-wdir" || a == "-wd") { need(1); wdir = argv[++i]; }
    else if (a == "-timeout") { need(1); timeout_s = atoi(argv[++i]); }
    else if (a == "-V" || a == "-version") { printf("mpirun (b200mpi) 0.1.0 — Open MPI/Hydra-compatible single-box launcher\n"); return 0; }
    else if (a == "-h" || a == "-help") {
      printf("usage: mpirun [options] <program> [args]      (aliases: mpiexec, mpiexec.hydra, orterun)\n"
             "Single-box launcher with the Open MPI and Hydra (MPICH / Intel MPI) command lines; one or two leading dashes.\n\n"
             "  -n, -np, -c N                 number of ranks (default: total slots of the hostfile, else 1)\n"
             "  -hostfile, -machinefile, -f F hostfile: 'host slots=N' (Open MPI) or 'host:N' (Hydra); default from\n"
             "                                OMPI_MCA_orte_default_hostfile, I_MPI_HYDRA_HOST_FILE, HYDRA_HOST_FILE\n"
             "  -host, -H, -hosts LIST        comma separated hosts (host[:slots])\n"
             "  -ppn, -perhost, -npernode N   ranks per host\n"
             "  -x VAR[=value]                export a variable to the ranks (Open MPI)\n"
             "  -genv, -env VAR value         export a variable to the ranks (Hydra); -genvall / -envall accepted\n"
             "  -mca / -gmca key value        exported as OMPI_MCA_<key>=<value>\n"
             "  -wdir, -wd DIR                working directory of the ranks\n"
             "  -tag-output, -prepend-rank, -l  prefix every output line with [job,rank]<stream>:\n"
             "  -output-filename, -output-directory DIR   also write every rank's output to DIR/1/rank.<N>/{stdout,stderr}\n"
             "  -merge-stderr-to-stdout       a rank's stderr travels with its stdout\n"
             "  -stdin RANK|all|none          which rank reads the launcher's stdin (default 0; the others get /dev/null)\n"
             "  -pernode                      one rank per host\n"
             "  -timestamp-output             prefix every forwarded line with the time\n"
             "  -timeout SECONDS              kill the job after SECONDS (exit code 124)\n"
             "  -oversubscribe                allow more ranks than slots\n"
             "  prog1 args : -np N prog2 ...  MPMD: several application contexts, ranks numbered context by context (MPI_APPNUM in\n"
             "                                OMPI_MCA_orte_app_num / PMI_APPNUM); a context may carry its own -np, -x, -env, -wdir\n"
             "  -bind-to none|numa|socket|core   processor binding (default none); numa = the CPUs next to the rank's GPU; with\n"
             "                   -cpus-per-proc N / -map-by X:PE=N, -cpu-set LIST, -report-bindings (B200MPI_BIND_TO overrides the default)\n"
             "  -map-by, -rank-by, -bootstrap, -launcher, -iface ...   accepted and ignored (one box, no ssh)\n"
             "  -V, -version / -h, -help\n\n"
             "Ranks get OMPI_COMM_WORLD_*, PMI_*, RANK/WORLD_SIZE/LOCAL_RANK/MASTER_ADDR/MASTER_PORT, HOROVOD_* and B200MPI_* variables;\n"
             "GPUs come from the operator's slot map (B200MPI_SLOTS_FILE). The first failing rank's exit code is propagated and the\n"
             "others are terminated (SIGTERM, then SIGKILL). B200MPI_FAULT=kill_rank:R@time:S injects a failure.\n");
      return 0;
    }
    else if (!a.empty() && a[0] == '-') { fprintf(stderr, "mpirun (b200mpi): note: ignoring unknown option %s\n", argv[i]); }
    else break;
  }
  if (i >= argc) die("no program to launch");
  // MPMD: `mpirun -np 1 ./master : -np 4 ./worker args` (Open MPI and Hydra): app contexts separated by a lone ':'; every
  // context after the first may start with its own -n/-np, -x / -env / -genv, -wdir; ranks are numbered context by context.
  struct App { int np = -1; std::vector<char*> argv; std::vector<std::pair<std::string, std::string>> env; std::string wdir; };
  std::vector<App> apps(1);
  apps[0].np = np;
  {
    bool opts = false;   // the first context's options were parsed above
    for (int k = i; k < argc; k++) {
      std::string a = argv[k];
      if (a == ":") {
        if (apps.back().argv.empty()) die("empty application context before ':'");
        apps.emplace_back();
        opts = true;
        continue;
      }
      if (opts && a.size() > 1 && a[0] == '-') {
        if (a.size() > 2 && a[1] == '-') a = a.substr(1);
        auto need_k = [&](int n) { if (k + n >= argc) die("option " + a + " needs an argument"); };
        if (a == "-n" || a == "-np" || a == "-c") { need_k(1); apps.back().np = atoi(argv[++k]); }
        else if (a == "-x") {
          need_k(1);
          std::string kv = argv[++k];
          size_t eq = kv.find('=');
          if (eq != std::string::npos) apps.back().env.push_back({kv.substr(0, eq), kv.substr(eq + 1)});
          else if (const char* v = getenv(kv.c_str())) apps.back().env.push_back({kv, v});
        }
        else if (a == "-env" || a == "-genv") { need_k(2); apps.back().env.push_back({argv[k + 1], argv[k + 2]}); k += 2; }
        else if (a == "-wdir" || a == "-wd") { need_k(1); apps.back().wdir = argv[++k]; }
        else if (a == "-host" || a == "-H" || a == "-hosts" || a == "-bind-to" || a == "-map-by" || a == "-mca") { need_k(a == "-mca" ? 2 : 1); k += a == "-mca" ? 2 : 1; }
        else fprintf(stderr, "mpirun (b200mpi): note: ignoring unknown option %s in application context %zu\n", argv[k], apps.size() - 1);
        continue;
      }
      opts = false;
      apps.back().argv.push_back(argv[k]);
    }
    if (apps.back().argv.empty()) die("no program after ':'");
    for (auto& ap : apps) ap.argv.push_back(nullptr);
    if (apps.size() > 1) {
      np = 0;
      for (auto& ap : apps) { if (ap.np <= 0) ap.np = 1; np += ap.np; }   // Open MPI: one process per context unless -np says otherwise
    }
  }
  auto app_of = [&](int rank) -> size_t {
    if (apps.size() == 1) return 0;
    int r = rank;
    for (size_t k = 0; k < apps.size(); k++) { if (r < apps[k].np) return k; r -= apps[k].np; }
    return apps.size() - 1;
  };

  // ---- hosts -------------------------------------------------------------
  int default_slots = 1;
  if (const char* s = getenv("OMPI_MCA_orte_set_default_slots")) default_slots = std::max(1, atoi(s));
  else if (const char* s2 = getenv("I_MPI_PERHOST")) default_slots = std::max(1, atoi(s2));
  if (ppn > 0) default_slots = ppn;
  if (hostfile.empty()) {
    for (const char* v : {"OMPI_MCA_orte_default_hostfile", "I_MPI_HYDRA_HOST_FILE", "HYDRA_HOST_FILE"})
      if (const char* p = getenv(v)) { if (access(p, R_OK) == 0) { hostfile = p; break; } }
  }
  std::vector<Host> hosts;
  if (!hostlist.empty()) {
    std::istringstream is(hostlist);
    std::string h;
    while (std::getline(is, h, ',')) {
      int slots = default_slots;
      size_t c = h.find(':');
      if (c != std::string::npos) { slots = atoi(h.c_str() + c + 1); h = h.substr(0, c); }
      hosts.push_back({h, slots});
    }
  } else if (!hostfile.empty()) {
    hosts = read_hostfile(hostfile, default_slots);
  }
  if (hosts.empty()) {
    char hn[256] = "localhost";
    gethostname(hn, sizeof(hn));
    hosts.push_back({hn, np > 0 ? np : default_slots});
  }
  if (ppn > 0) for (auto& h : hosts) h.slots = ppn;
  int total_slots = 0;
  for (auto& h : hosts) total_slots += h.slots;
  if (np <= 0) np = total_slots;
  if (pe_given && !bind_explicit) bind_policy = BIND_CORE;
  if (np > total_slots && !oversubscribe && getenv("B200MPI_STRICT_SLOTS"))
    die("not enough slots: requested " + std::to_string(np) + ", hostfile provides " + std::to_string(total_slots));

  // ---- placement by slot -----------------------------------------------------
  std::vector<Rank> ranks(np);
  {
    int r = 0;
    std::vector<int> used(hosts.size(), 0);
    while (r < np) {
      bool placed = false;
      for (size_t h = 0; h < hosts.size() && r < np; h++) {
        int cap = hosts[h].slots;
        int start = used[h];
        // fill this host's remaining slots (first pass), or one more round when oversubscribed
        for (int s = start; s < ((start / cap) + 1) * cap && r < np; s++) {
          ranks[r].rank = r; ranks[r].host = hosts[h].name; ranks[r].node = (int)h; ranks[r].local_rank = s;
          used[h]++; r++; placed = true;
        }
      }
      if (!placed) break;
    }
  }
  std::vector<int> local_size(hosts.size(), 0);
  for (auto& rk : ranks) local_size[rk.node]++;

  // ---- GPU map from the node agent ---------------------------------------------
  // Like the reference's entrypoint.sh waiting for DNS (build/base/entrypoint.sh:8-34): give the node
  // agent a moment to publish every host of the hostfile in the slot map (elastic scale-up races).
  auto slots = read_slots_file(getenv("B200MPI_SLOTS_FILE"));
  if (getenv("B200MPI_SLOTS_FILE") && !slots.empty()) {
    for (int attempt = 0; attempt < 50; attempt++) {
      bool all = true;
      for (auto& h : hosts) all = all && slots.count(short_host(h.name)) > 0;
      if (all) break;
      usleep(100000);
      slots = read_slots_file(getenv("B200MPI_SLOTS_FILE"));
    }
  }
  std::vector<int> job_gpus;
  for (auto& h : hosts) {
    auto it = slots.find(short_host(h.name));
    if (it != slots.end()) job_gpus.insert(job_gpus.end(), it->second.begin(), it->second.end());
  }
  std::string cvd;
  if ((int)job_gpus.size() >= np) {
    for (int k = 0; k < np; k++) cvd += (k ? "," : "") + std::to_string(job_gpus[k]);
  }

  // GPU of every rank as far as binding is concerned: the node agent's reservation, else the launcher's own CUDA_VISIBLE_DEVICES
  // (numeric entries), else rank r -> GPU r (the LOCAL_RANK convention of every workload here)
  auto binding_gpus = [&]() {
    if (!cvd.empty()) return job_gpus;
    std::vector<int> g;
    const char* env = getenv("CUDA_VISIBLE_DEVICES");
    if (env && *env) {
      std::istringstream is(env);
      std::string tok;
      while (std::getline(is, tok, ',')) g.push_back(!tok.empty() && isdigit((unsigned char)tok[0]) ? atoi(tok.c_str()) : -1);
    } else {
      for (int r = 0; r < np; r++) g.push_back(r);
    }
    return g;
  };

  // ---- rendezvous identity -------------------------------------------------------
  std::string job_id = getenv("B200MPI_JOB_ID") ? getenv("B200MPI_JOB_ID") : "mpirun";
  job_id += "." + std::to_string((long)getpid());
  unsigned hash = 5381;
  for (char c : job_id) hash = hash * 33u + (unsigned char)c;
  const int master_port = 20000 + (int)(hash % 20000u);

  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  signal(SIGHUP, on_signal);
  signal(SIGPIPE, SIG_IGN);
  if (!wdir.empty() && chdir(wdir.c_str()) != 0) die("cannot chdir to " + wdir);

  // On one box the whole job is a single NVSwitch node: LOCAL_RANK == RANK indexes
  // the job's GPU list.
  const bool single_node_view = !cvd.empty() || getenv("B200MPI_SINGLE_NODE_VIEW");
  // ---- elastic rescale in place (B200MPI_ELASTIC_INPLACE=1): this launcher watches discover_hosts.sh itself; see the
  // supervise loop. Ranks learn the world from <elastic_dir>/world = "<generation> <world size>".
  const bool elastic = getenv("B200MPI_ELASTIC_INPLACE") && atoi(getenv("B200MPI_ELASTIC_INPLACE")) != 0 && apps.size() == 1;
  std::string elastic_dir, discover_path;
  int generation = 0, world_now = np;
  if (elastic) {
    elastic_dir = std::string(getenv("B200MPI_POD_DIR") ? getenv("B200MPI_POD_DIR") : "/tmp") + "/b200mpi-elastic-" + std::to_string((long)getpid());
    mkdir(elastic_dir.c_str(), 0755);
    if (const char* dp = getenv("B200MPI_DISCOVER_HOSTS")) discover_path = dp;
    else {
      std::string root = getenv("B200MPI_POD_ROOTFS") ? getenv("B200MPI_POD_ROOTFS") : "";
      discover_path = root + "/etc/mpi/discover_hosts.sh";
      if (access(discover_path.c_str(), R_OK) != 0) discover_path = "/etc/mpi/discover_hosts.sh";
    }
  }
  auto publish_world = [&](int gen, int world) {
    const std::string tmp = elastic_dir + "/world.tmp", fin = elastic_dir + "/world";
    if (FILE* f = fopen(tmp.c_str(), "w")) { fprintf(f, "%d %d\n", gen, world); fclose(f); rename(tmp.c_str(), fin.c_str()); }
  };
  if (elastic) publish_world(0, np);

  // spawn one rank into a world of `world` ranks (generation `gen`; join: it enters a RUNNING elastic job)
  auto spawn_rank = [&](Rank& rk, int world, int gen, bool join) {
    const int np = world;   // everything below describes the world this rank is born into
    int po[2], pe[2];
    if (pipe2(po, O_CLOEXEC) || pipe2(pe, O_CLOEXEC)) die(std::string("pipe: ") + strerror(errno));
    pid_t pid = fork();
    if (pid < 0) die(std::string("fork: ") + strerror(errno));
    if (pid == 0) {
      setpgid(0, 0);
      dup2(po[1], 1);
      dup2(g_merge_stderr ? po[1] : pe[1], 2);
      if (g_stdin_rank != -2 && g_stdin_rank != rk.rank) {   // Open MPI: only rank 0 (or -stdin RANK) reads the launcher's stdin
        const int nul = open("/dev/null", O_RDONLY);
        if (nul >= 0) { dup2(nul, 0); if (nul > 2) close(nul); }
      }
      signal(SIGPIPE, SIG_DFL);
      signal(SIGINT, SIG_DFL);
      signal(SIGTERM, SIG_DFL);
      const int lrank = single_node_view ? rk.rank : rk.local_rank;
      const int lsize = single_node_view ? np : local_size[std::min((size_t)rk.node, local_size.size() - 1)];
      const int node = single_node_view ? 0 : rk.node;
      auto set = [](const char* k, const std::string& v) { setenv(k, v.c_str(), 1); };
      auto seti = [&](const char* k, int v) { set(k, std::to_string(v)); };
      for (auto& kv : xenv) set(kv.first.c_str(), kv.second);
      set("B200MPI_JOB_ID", job_id);
      if (elastic) {
        set("B200MPI_ELASTIC_DIR", elastic_dir);
        seti("B200MPI_GENERATION", gen);
        if (join) set("B200MPI_ELASTIC_JOIN", "1");
      }
      seti("B200MPI_RANK", rk.rank); seti("B200MPI_WORLD_SIZE", np);
      seti("B200MPI_LOCAL_RANK", lrank); seti("B200MPI_LOCAL_SIZE", lsize);
      seti("OMPI_COMM_WORLD_RANK", rk.rank); seti("OMPI_COMM_WORLD_SIZE", np);
      seti("OMPI_COMM_WORLD_LOCAL_RANK", lrank); seti("OMPI_COMM_WORLD_LOCAL_SIZE", lsize);
      seti("OMPI_COMM_WORLD_NODE_RANK", lrank); seti("OMPI_UNIVERSE_SIZE", total_slots);
      seti("PMI_RANK", rk.rank); seti("PMI_SIZE", np); seti("MPI_LOCALRANKID", lrank); seti("MPI_LOCALNRANKS", lsize);
      seti("PMIX_RANK", rk.rank);
      seti("RANK", rk.rank); seti("WORLD_SIZE", np); seti("LOCAL_RANK", lrank); seti("LOCAL_WORLD_SIZE", lsize);
      seti("GROUP_RANK", node); set("MASTER_ADDR", "127.0.0.1"); seti("MASTER_PORT", master_port);
      seti("HOROVOD_RANK", rk.rank); seti("HOROVOD_SIZE", np); seti("HOROVOD_LOCAL_RANK", lrank);
      seti("HOROVOD_LOCAL_SIZE", lsize); seti("HOROVOD_CROSS_RANK", node); set("HOROVOD_CROSS_SIZE", "1");
      set("HOROVOD_HOSTNAME", rk.host);
      set("K_MPI_JOB_ROLE", "worker");
      set("B200MPI_HOSTNAME", short_host(rk.host));
      set("HOSTNAME", short_host(rk.host));
      if (!cvd.empty() && rk.rank < (int)job_gpus.size()) {
        set("CUDA_VISIBLE_DEVICES", cvd);
        seti("B200MPI_GPU", job_gpus[rk.rank]);
        unsetenv("NVIDIA_VISIBLE_DEVICES");
        unsetenv("NVIDIA_DRIVER_CAPABILITIES");
      } else if (getenv("CUDA_VISIBLE_DEVICES") && !*getenv("CUDA_VISIBLE_DEVICES") && getenv("B200MPI_GPUS") && *getenv("B200MPI_GPUS")) {
        set("CUDA_VISIBLE_DEVICES", getenv("B200MPI_GPUS"));
      }
      if (const char* inj = getenv("B200MPI_INJECT_LIB")) {  // LD-inject the collective runtime
        if (*inj && access(inj, R_OK) == 0) {
          std::string pre = inj;
          if (const char* old = getenv("LD_PRELOAD")) if (*old) pre += std::string(":") + old;
          set("LD_PRELOAD", pre);
        }
      }
      if (bind_policy != BIND_NONE) {
        int numa = -1;
        const std::vector<int> cpus = binding_for(bind_policy, rk.rank, np, binding_gpus(), cpus_per_proc, cpu_set, &numa);
        if (!cpus.empty()) {
          cpu_set_t m;
          CPU_ZERO(&m);
          for (int c : cpus) CPU_SET(c, &m);
          if (sched_setaffinity(0, sizeof(m), &m) == 0) { set("B200MPI_BOUND_CPUS", format_cpulist(cpus)); seti("B200MPI_BOUND_NUMA", numa); }
        }
      }
      const size_t appnum = app_of(rk.rank);
      App& ap = apps[appnum];
      for (auto& kv : ap.env) set(kv.first.c_str(), kv.second);
      seti("OMPI_MCA_orte_app_num", (int)appnum); seti("PMI_APPNUM", (int)appnum); seti("B200MPI_APPNUM", (int)appnum);   // MPI_APPNUM
      if (!ap.wdir.empty() && chdir(ap.wdir.c_str()) != 0) { fprintf(stderr, "mpirun (b200mpi): cannot chdir to %s\n", ap.wdir.c_str()); _exit(127); }
      char** prog = ap.argv.data();
      execvp(prog[0], prog);
      if (errno == ENOENT && strchr(prog[0], '/')) {
        // image-relative path (e.g. /home/mpiuser/pi from the reference YAML): fall back to PATH lookup
        const char* base = strrchr(prog[0], '/') + 1;
        execvp(base, prog);
      }
      fprintf(stderr, "mpirun (b200mpi): could not exec %s: %s\n", prog[0], strerror(errno));
      _exit(127);
    }
    setpgid(pid, pid);
    if (report_bindings) {
      int numa = -1;
      const std::vector<int> cpus = binding_for(bind_policy, rk.rank, np, binding_gpus(), cpus_per_proc, cpu_set, &numa);
      if (cpus.empty()) fprintf(stderr, "[%s:%d] MCW rank %d is not bound (or bound to all available processors)\n", short_host(rk.host).c_str(), (int)pid, rk.rank);
      else fprintf(stderr, "[%s:%d] MCW rank %d bound to NUMA node %d: cpus %s%s\n", short_host(rk.host).c_str(), (int)pid, rk.rank, numa,
                   format_cpulist(cpus).c_str(), (!cvd.empty() && rk.rank < (int)job_gpus.size()) ? (" (GPU " + std::to_string(job_gpus[rk.rank]) + ")").c_str() : "");
      fflush(stderr);
    }
    close(po[1]); close(pe[1]);
    fcntl(po[0], F_SETFL, O_NONBLOCK);
    fcntl(pe[0], F_SETFL, O_NONBLOCK);
    rk.pid = pid; rk.out = po[0]; rk.err = pe[0];
    rk.pidfd = (int)syscall(SYS_pidfd_open, pid, 0);  // -1 on old kernels: waitpid polling below still works
  };
  for (auto& rk : ranks) spawn_rank(rk, np, 0, false);

  // ---- fault injection from outside the ranks (mpi_operator_b200/utils/fault.py documents the grammar):
  // B200MPI_FAULT="kill_rank:R@time:SECONDS[;once]" — SIGKILL rank R's process group SECONDS after spawn.
  int fault_rank = -1; double fault_after = 0; bool fault_done = false; std::string fault_marker;
  if (const char* fs = getenv("B200MPI_FAULT")) {
    int r = -1; double t = 0;
    if (sscanf(fs, "kill_rank:%d@time:%lf", &r, &t) == 2 && r >= 0 && r < np) {
      fault_rank = r; fault_after = t;
      if (strstr(fs, ";once")) {
        std::string dir = getenv("B200MPI_FAULT_DIR") ? getenv("B200MPI_FAULT_DIR") : "";
        if (dir.empty() && getenv("B200MPI_SLOTS_FILE")) { dir = getenv("B200MPI_SLOTS_FILE"); size_t k = dir.rfind('/'); dir = k == std::string::npos ? "." : dir.substr(0, k); }
        if (dir.empty()) dir = "/tmp";
        fault_marker = dir + "/" + (getenv("B200MPI_MPIJOB_NAME") ? getenv("B200MPI_MPIJOB_NAME") : "job") + ".fault-mpirun.fired";
        if (access(fault_marker.c_str(), F_OK) == 0) fault_done = true;
      }
    }
  }
  struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);

  // ---- supervise --------------------------------------------------------------------
  int alive = np, first_fail_rank = -1, first_fail_status = 0;
  bool killing = false;
  time_t t_start = time(nullptr), t_kill = 0;
  auto kill_all = [&](int sig) {
    for (auto& rk : ranks) if (!rk.exited && rk.pid > 0) kill(-rk.pid, sig);
  };
  // elastic state: `target` = world size discover_hosts.sh currently describes; a grow first spawns the new ranks, waits for
  // their ready files (python + torch imported, about to join) and only then publishes the world, so the survivors keep
  // training on the old world while the newcomers boot.
  int el_target = np, el_stable = 0, el_pending_world = 0;
  time_t el_pending_since = 0;
  struct timespec el_last = ts0;
  auto discover_world = [&]() -> int {
    FILE* f = fopen(discover_path.c_str(), "r");
    if (!f) return -1;
    char line[1024];
    int hosts_n = 0;
    while (fgets(line, sizeof(line), f)) if (strncmp(line, "echo ", 5) == 0) hosts_n++;
    fclose(f);
    return hosts_n * default_slots;
  };
  while (alive > 0 || std::any_of(ranks.begin(), ranks.end(), [](const Rank& r) { return r.out >= 0 || r.err >= 0; })) {
    if (elastic && !killing && alive > 0) {
      struct timespec tn; clock_gettime(CLOCK_MONOTONIC, &tn);
      if ((tn.tv_sec - el_last.tv_sec) + (tn.tv_nsec - el_last.tv_nsec) * 1e-9 >= 0.25) {
        el_last = tn;
        if (el_pending_world == 0) {
          const int w = discover_world();
          if (w > 0 && w != world_now) {
            if (w == el_target) el_stable++; else { el_target = w; el_stable = 0; }
            if (el_stable >= 2) {   // the same new world for three consecutive reads (the controller rewrites the file pod by pod)
              generation++;
              el_pending_world = w;
              el_pending_since = time(nullptr);
              if (w > world_now) {
                // GPU map of the larger world, then the additional ranks
                auto slots2 = read_slots_file(getenv("B200MPI_SLOTS_FILE"));
                if (!slots2.empty() && !cvd.empty()) {
                  std::vector<int> g2;
                  std::vector<std::string> names;
                  for (auto& kv : slots2) names.push_back(kv.first);
                  // keep the survivors' order: known hosts first (hostfile order), new hosts after them in name order
                  for (auto& h : hosts) { auto it = slots2.find(short_host(h.name)); if (it != slots2.end()) g2.insert(g2.end(), it->second.begin(), it->second.end()); }
                  for (auto& nme : names) {
                    bool known = false;
                    for (auto& h : hosts) known = known || short_host(h.name) == nme;
                    if (!known) { g2.insert(g2.end(), slots2[nme].begin(), slots2[nme].end()); hosts.push_back({nme, default_slots}); }
                  }
                  if ((int)g2.size() >= w) {
                    job_gpus = g2;
                    cvd.clear();
                    for (int k = 0; k < w; k++) cvd += (k ? "," : "") + std::to_string(job_gpus[k]);
                  }
                }
                const int first_new = (int)ranks.size();
                for (int r = first_new; r < w; r++) {
                  Rank nr;
                  nr.rank = r; nr.local_rank = r; nr.node = 0; nr.host = r < (int)hosts.size() ? hosts[r].name : (hosts.empty() ? "localhost" : hosts.back().name);
                  ranks.push_back(nr);
                }
                for (int r = first_new; r < w; r++) { spawn_rank(ranks[r], w, generation, true); alive++; }
                fprintf(stderr, "mpirun (b200mpi): elastic: world %d -> %d, generation %d: spawned %d ranks, waiting until they are ready\n",
                        world_now, w, generation, w - first_new);
              } else {
                for (auto& rk : ranks) if (rk.rank >= w) rk.retired = true;
                publish_world(generation, w);
                fprintf(stderr, "mpirun (b200mpi): elastic: world %d -> %d, generation %d: ranks >= %d retire at their next commit\n",
                        world_now, w, generation, w);
                world_now = w;
                el_pending_world = 0;
              }
            }
          } else if (w == world_now) { el_target = w; el_stable = 0; }
        } else {
          // grow in progress: publish once every newcomer has announced itself (or after 120 s regardless)
          bool all_ready = true;
          for (int r = world_now; r < el_pending_world; r++) {
            const std::string rf = elastic_dir + "/ready." + std::to_string(generation) + "." + std::to_string(r);
            all_ready = all_ready && access(rf.c_str(), F_OK) == 0;
          }
          if (all_ready || time(nullptr) - el_pending_since > 120) {
            publish_world(generation, el_pending_world);
            fprintf(stderr, "mpirun (b200mpi): elastic: generation %d published (%d ranks)\n", generation, el_pending_world);
            world_now = el_pending_world;
            el_pending_world = 0;
          }
        }
      }
    }
    std::vector<pollfd> pf;
    for (auto& rk : ranks) {
      if (rk.out >= 0) pf.push_back({rk.out, POLLIN, 0});
      if (rk.err >= 0) pf.push_back({rk.err, POLLIN, 0});
      if (!rk.exited && rk.pidfd >= 0) pf.push_back({rk.pidfd, POLLIN, 0});
    }
    poll(pf.data(), pf.size(), 100);
    if (fault_rank >= 0 && !fault_done) {
      struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
      if ((ts.tv_sec - ts0.tv_sec) + (ts.tv_nsec - ts0.tv_nsec) * 1e-9 >= fault_after) {
        fault_done = true;
        if (!fault_marker.empty()) { FILE* mf = fopen(fault_marker.c_str(), "w"); if (mf) { fprintf(mf, "rank %d\n", fault_rank); fclose(mf); } }
        for (auto& rk : ranks) if (rk.rank == fault_rank && !rk.exited && rk.pid > 0) {
          fprintf(stderr, "mpirun (b200mpi): fault injection: SIGKILL rank %d\n", fault_rank);
          kill(-rk.pid, SIGKILL);
        }
      }
    }
    for (auto& rk : ranks) {
      if (rk.out >= 0) { bool open = drain(rk.out, rk.obuf); flush_lines(rk.obuf, stdout, tag_output, rk.rank, "stdout", !open); }
      if (rk.err >= 0) { bool open = drain(rk.err, rk.ebuf); flush_lines(rk.ebuf, stderr, tag_output, rk.rank, "stderr", !open); }
      if (!rk.exited) {
        int st = 0;
        pid_t w = waitpid(rk.pid, &st, WNOHANG);
        if (w == rk.pid) {
          rk.exited = true; rk.status = st; alive--;
          if (rk.pidfd >= 0) { close(rk.pidfd); rk.pidfd = -1; }
          bool bad = !(WIFEXITED(st) && WEXITSTATUS(st) == 0);
          if (rk.retired && WIFEXITED(st) && WEXITSTATUS(st) == 75) bad = false;   // left with the rescale code after a scale-down
          if (bad && first_fail_rank < 0 && !killing) { first_fail_rank = rk.rank; first_fail_status = st; }
        }
      }
    }
    if ((first_fail_rank >= 0 || g_signal || (timeout_s > 0 && time(nullptr) - t_start > timeout_s)) && !killing && alive > 0) {
      killing = true; t_kill = time(nullptr);
      kill_all(g_signal ? (int)g_signal : SIGTERM);
    }
    if (killing && alive > 0 && time(nullptr) - t_kill >= 3) kill_all(SIGKILL);
    if (alive == 0) {
      // children may leave grandchildren holding the pipes: don't wait for them forever
      bool pending = false;
      for (auto& rk : ranks) pending = pending || rk.out >= 0 || rk.err >= 0;
      if (pending) {
        usleep(50000);
        for (auto& rk : ranks) {
          if (rk.out >= 0) { drain(rk.out, rk.obuf); flush_lines(rk.obuf, stdout, tag_output, rk.rank, "stdout", true); if (rk.out >= 0) { close(rk.out); rk.out = -1; } }
          if (rk.err >= 0) { drain(rk.err, rk.ebuf); flush_lines(rk.ebuf, stderr, tag_output, rk.rank, "stderr", true); if (rk.err >= 0) { close(rk.err); rk.err = -1; } }
        }
      }
    }
  }
  if (g_signal) return 128 + (int)g_signal;
  if (timeout_s > 0 && killing && first_fail_rank < 0) { fprintf(stderr, "mpirun (b200mpi): job exceeded --timeout %d s\n", timeout_s); return 124; }
  // Every rank is gone: remove the job's shared-memory segments (runtime, libmpi shim, Horovod-core engine). The ranks unlink
  // them themselves on a clean shutdown; after a crash, a kill or a script that simply exits they would stay in /dev/shm forever.
  {
    std::string key;
    for (char ch : job_id) key.push_back((isalnum((unsigned char)ch) || ch == '-' || ch == '_' || ch == '.') ? ch : '_');
    if (DIR* d = opendir("/dev/shm")) {
      while (dirent* de = readdir(d)) {
        const std::string n = de->d_name;
        // names are b200mpi-<id>[-suffix] and b200mpi-mpi-<id>[-suffix]: match the id exactly, not as a substring (pid 123 vs 1234)
        for (const char* prefix : {"b200mpi-mpi-", "b200mpi-"}) {
          const size_t pl = strlen(prefix);
          if (n.compare(0, pl, prefix) != 0) continue;
          const std::string rest = n.substr(pl);
          if (rest == key || rest.compare(0, key.size() + 1, key + "-") == 0) { unlink(("/dev/shm/" + n).c_str()); break; }
        }
      }
      closedir(d);
    }
  }
  if (first_fail_rank >= 0) {
    int code = WIFEXITED(first_fail_status) ? WEXITSTATUS(first_fail_status) : 128 + WTERMSIG(first_fail_status);
    fprintf(stderr,
            "--------------------------------------------------------------------------\n"
            "mpirun (b200mpi) detected that one or more processes exited with non-zero status,\n"
            "thus causing the job to be terminated. The first process to do so was:\n\n"
            "  Process name: [[b200mpi],%d]\n  Exit code:    %d\n"
            "--------------------------------------------------------------------------\n",
            first_fail_rank, code);
    return code ? code : 1;
  }
  return 0;
}
