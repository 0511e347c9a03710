/*
 * ohp_boot.cpp -- MPI-free start-up of the one-process-per-GPU job.
 *
 * The reference gets rank, size and a communicator from PetscInitialize -> MPI_Init (ex12.cpp:1485,
 * MPI_Comm_rank/size at ex12.cpp:800-801).  This image has no MPI, and the launchers in use
 * (torch.distributed.run, or any wrapper that exports RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
 * MASTER_PORT) only hand out environment variables.  The single thing the ranks must share before NCCL can
 * take over is the 128-byte ncclUniqueId: rank 0 serves it on a TCP port next to MASTER_PORT, the other ranks
 * connect (with retries while rank 0 is still starting) and read it.  Everything after that -- cell counts,
 * CUDA IPC handles, gid round trips -- travels over the NCCL communicator itself.
 */
#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdlib.h>
#include <string.h>
#include <sys/select.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#include "ohp_internal.h"

namespace {

int env_int(const char *a, const char *b, const char *c, int dflt) {
  const char *names[3] = {a, b, c};
  for (const char *n : names) {
    if (!n) continue;
    const char *v = getenv(n);
    if (v && *v) return atoi(v);
  }
  return dflt;
}

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int write_all(int fd, const void *buf, size_t n) {
  const char *p = (const char *)buf;
  while (n) {
    const ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; return -1; }
    p += k; n -= (size_t)k;
  }
  return 0;
}
int read_all(int fd, void *buf, size_t n) {
  char *p = (char *)buf;
  while (n) {
    const ssize_t k = recv(fd, p, n, 0);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; return -1; }
    p += k; n -= (size_t)k;
  }
  return 0;
}

int boot_port() {
  const char *p = getenv("OHP_BOOT_PORT");
  if (p && *p) return atoi(p);
  return env_int("MASTER_PORT", nullptr, nullptr, 29500) + 307;
}
double boot_timeout() {
  const char *t = getenv("OHP_BOOT_TIMEOUT_S");
  const double s = t ? atof(t) : 120.0;
  return s > 0.0 ? s : 120.0;
}

}  // namespace

/* MPI_Comm_rank / MPI_Comm_size (ex12.cpp:800-801) from the launcher's environment:
 * torch.distributed.run (RANK, WORLD_SIZE, LOCAL_RANK), Open MPI / MPICH / Slurm names as fall-backs */
extern "C" int ohp_boot_env(int *rank, int *nranks, int *local_rank) {
  const int r = env_int("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", env_int("SLURM_PROCID", nullptr, nullptr, 0));
  const int n = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", env_int("SLURM_NTASKS", nullptr, nullptr, 1));
  const int l = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", env_int("SLURM_LOCALID", nullptr, nullptr, r));
  if (n < 1 || r < 0 || r >= n) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_boot_env: rank %d of %d", r, n);
  if (rank) *rank = r;
  if (nranks) *nranks = n;
  if (local_rank) *local_rank = l;
  return 0;
}

/* Broadcast `bytes` bytes from rank 0 to every rank over TCP (MASTER_ADDR : OHP_BOOT_PORT or MASTER_PORT + 307).
 * Plays the role of the MPI_Bcast a maintainer would use for the ncclUniqueId. */
extern "C" int ohp_boot_bcast(int rank, int nranks, void *buf, int bytes) {
  if (nranks <= 1) return 0;
  if (!buf || bytes <= 0) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_boot_bcast: empty buffer");
  const int port = boot_port();
  const double deadline = now_s() + boot_timeout();
  if (rank == 0) {
    const int srv = socket(AF_INET, SOCK_STREAM, 0);
    if (srv < 0) OHP_FAIL(OHP_ERR_LIB, "ohp_boot_bcast: socket(): %s", strerror(errno));
    int one = 1;
    setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
    addr.sin_port = htons((uint16_t)port);
    while (bind(srv, (sockaddr *)&addr, sizeof(addr)) != 0) {
      if (now_s() > deadline) { close(srv); OHP_FAIL(OHP_ERR_LIB, "ohp_boot_bcast: cannot bind port %d: %s (set OHP_BOOT_PORT)", port, strerror(errno)); }
      usleep(100000);
    }
    if (listen(srv, nranks) != 0) { close(srv); OHP_FAIL(OHP_ERR_LIB, "ohp_boot_bcast: listen(): %s", strerror(errno)); }
    std::vector<char> seen(nranks, 0);
    int served = 0;
    while (served < nranks - 1) {
      timeval tv;
      const double left = deadline - now_s();
      if (left <= 0.0) { close(srv); OHP_FAIL(OHP_ERR_LIB, "ohp_boot_bcast: only %d of %d ranks connected within the time limit", served, nranks - 1); }
      tv.tv_sec = (long)left; tv.tv_usec = 0;
      fd_set fds;
      FD_ZERO(&fds);
      FD_SET(srv, &fds);
      if (select(srv + 1, &fds, nullptr, nullptr, &tv) <= 0) continue;
      const int c = accept(srv, nullptr, nullptr);
      if (c < 0) continue;
      int32_t hello[2] = {-1, -1};
      if (read_all(c, hello, sizeof(hello)) == 0 && hello[1] == nranks && hello[0] > 0 && hello[0] < nranks && !seen[hello[0]] &&
          write_all(c, buf, (size_t)bytes) == 0) {
        seen[hello[0]] = 1;
        served++;
      }
      close(c);
    }
    close(srv);
    return 0;
  }
  const char *host = getenv("MASTER_ADDR");
  if (!host || !*host) host = "127.0.0.1";
  addrinfo hints, *res = nullptr;
  memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  char portstr[16];
  snprintf(portstr, sizeof(portstr), "%d", port);
  if (getaddrinfo(host, portstr, &hints, &res) != 0 || !res) OHP_FAIL(OHP_ERR_LIB, "ohp_boot_bcast: cannot resolve MASTER_ADDR %s", host);
  for (;;) {
    const int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd >= 0 && connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
      int one = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      const int32_t hello[2] = {rank, nranks};
      const bool ok = write_all(fd, hello, sizeof(hello)) == 0 && read_all(fd, buf, (size_t)bytes) == 0;
      close(fd);
      if (ok) { freeaddrinfo(res); return 0; }
    } else if (fd >= 0) close(fd);
    if (now_s() > deadline) { freeaddrinfo(res); OHP_FAIL(OHP_ERR_LIB, "ohp_boot_bcast: rank %d could not reach rank 0 at %s:%d", rank, host, port); }
    usleep(50000);
  }
}

/* PetscInitialize's communicator for a multi-rank launch: rank/size from the environment, the NCCL id from rank 0 */
extern "C" int ohp_comm_init_from_env(ohp_context ctx) {
  if (!ctx) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_comm_init_from_env: null context");
  int rank = 0, nranks = 1;
  OHP_TRY(ohp_boot_env(&rank, &nranks, nullptr));
  if (nranks == 1) return 0;
  unsigned char id[128];
  memset(id, 0, sizeof(id));
  if (rank == 0) OHP_TRY(ohp_comm_unique_id(id));
  OHP_TRY(ohp_boot_bcast(rank, nranks, id, (int)sizeof(id)));
  return ohp_comm_init(ctx, nranks, rank, id);
}
