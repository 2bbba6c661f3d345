/* nccl_emu.c -- TEST-ONLY stand-in for the handful of NCCL entry points libathena_b200 loads with
 * dlopen, for the CPU test of the multi-rank path: several PROCESSES, each running the emulated
 * device path (tests/hostcheck/libathena_b200_emu.so, "device" memory = host memory), exchange
 * their ghost zones, EMF corrections and dt minima through UNIX-domain sockets.
 *
 *   ncclGetUniqueId     a socket-path prefix in the 128-byte id
 *   ncclCommInitRank    every pair of ranks gets one stream socket (the lower rank listens)
 *   ncclSend / ncclRecv inside ncclGroupStart/End: queued, then progressed together with poll()
 *                       (never blocks on one peer while another waits); outside a group: at once.
 *                       Matching is by order and size per pair, as in NCCL.
 *   ncclAllReduce       doubles, MIN or SUM: gathered on rank 0 in rank order, result sent back
 *
 * The product never links or loads this file: tests point the loader at it with AB_NCCL_LIB.
 * Build: gcc -O1 -g -fPIC -shared -o libnccl_emu.so nccl_emu.c */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

typedef struct { char internal[128]; } ncclUniqueId;
typedef struct Comm {
  int nranks, rank;
  int *fd;                 /* socket to every peer (-1 for self) */
  char path[160];
} Comm;
typedef Comm *ncclComm_t;

/* every message is preceded by its byte count: a receive that expects another size fails
 * loudly (real NCCL would hang or corrupt memory) */
typedef struct { int send; char *p; size_t left; int peer; Comm *c; size_t size; size_t hdr; int hleft; } Op;
static Op g_ops[4096];
static int g_nops = 0, g_depth = 0;

const char *ncclGetErrorString(int r) { return r ? "nccl_emu error" : "no error"; }

int ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof(*id));
  const char *tmp = getenv("TMPDIR");
  snprintf(id->internal, sizeof(id->internal), "%s/abnccl_%d_%ld", tmp ? tmp : "/tmp", (int)getpid(),
           (long)time(NULL) ^ (long)(size_t)id);
  return 0;
}

static int xwrite(int fd, const void *p, size_t n) {
  const char *q = (const char *)p;
  while (n) { ssize_t k = write(fd, q, n); if (k < 0) { if (errno == EINTR) continue; return 1; } q += k; n -= k; }
  return 0;
}
static int xread(int fd, void *p, size_t n) {
  char *q = (char *)p;
  while (n) { ssize_t k = read(fd, q, n); if (k <= 0) { if (k < 0 && errno == EINTR) continue; return 1; } q += k; n -= k; }
  return 0;
}

int ncclCommInitRank(ncclComm_t *out, int nranks, ncclUniqueId id, int rank) {
  Comm *c = (Comm *)calloc(1, sizeof(Comm));
  c->nranks = nranks; c->rank = rank;
  c->fd = (int *)malloc(sizeof(int)*nranks);
  for (int r = 0; r < nranks; ++r) c->fd[r] = -1;
  snprintf(c->path, sizeof(c->path), "%s.%d", id.internal, rank);
  int ls = -1;
  if (rank < nranks - 1) {            /* ranks above me connect to me */
    ls = socket(AF_UNIX, SOCK_STREAM, 0);
    struct sockaddr_un a; memset(&a, 0, sizeof(a)); a.sun_family = AF_UNIX;
    strncpy(a.sun_path, c->path, sizeof(a.sun_path) - 1);
    unlink(c->path);
    if (bind(ls, (struct sockaddr *)&a, sizeof(a)) || listen(ls, nranks)) return 2;
  }
  for (int r = 0; r < rank; ++r) {    /* connect to every lower rank (retry until it listens) */
    struct sockaddr_un a; memset(&a, 0, sizeof(a)); a.sun_family = AF_UNIX;
    snprintf(a.sun_path, sizeof(a.sun_path), "%s.%d", id.internal, r);
    int s = -1;
    for (int tries = 0; tries < 60000; ++tries) {
      s = socket(AF_UNIX, SOCK_STREAM, 0);
      if (connect(s, (struct sockaddr *)&a, sizeof(a)) == 0) break;
      close(s); s = -1;
      usleep(1000);
    }
    if (s < 0) return 3;
    int me = rank;
    if (xwrite(s, &me, sizeof(me))) return 4;
    c->fd[r] = s;
  }
  for (int k = rank + 1; k < nranks; ++k) {
    int s = accept(ls, NULL, NULL);
    int who = -1;
    if (s < 0 || xread(s, &who, sizeof(who)) || who <= rank || who >= nranks) return 5;
    c->fd[who] = s;
  }
  if (ls >= 0) { close(ls); unlink(c->path); }
  *out = c;
  return 0;
}

int ncclCommDestroy(ncclComm_t c) {
  if (!c) return 0;
  for (int r = 0; r < c->nranks; ++r) if (c->fd[r] >= 0) close(c->fd[r]);
  free(c->fd); free(c);
  return 0;
}

static int progress_all(void) {
  /* every queued op on its socket, in order per (peer, direction), all peers concurrently */
  for (int i = 0; i < g_nops; ++i) {
    int fl = fcntl(g_ops[i].c->fd[g_ops[i].peer], F_GETFL, 0);
    fcntl(g_ops[i].c->fd[g_ops[i].peer], F_SETFL, fl | O_NONBLOCK);
  }
  int live = g_nops, rc = 0;
  while (live > 0 && !rc) {
    struct pollfd pf[4096];
    int idx[4096], n = 0;
    /* per (fd, direction) only the FIRST unfinished op may move: order is the matching rule */
    for (int i = 0; i < g_nops; ++i) {
      if (!g_ops[i].left) continue;
      int fd = g_ops[i].c->fd[g_ops[i].peer], dup = 0;
      for (int j = 0; j < n; ++j)
        if (pf[j].fd == fd && ((pf[j].events & POLLOUT) != 0) == (g_ops[i].send != 0)) { dup = 1; break; }
      if (dup) continue;
      pf[n].fd = fd; pf[n].events = g_ops[i].send ? POLLOUT : POLLIN; pf[n].revents = 0; idx[n] = i; ++n;
    }
    if (poll(pf, n, 60000) <= 0) { rc = 6; break; }
    for (int j = 0; j < n; ++j) {
      if (!(pf[j].revents & (POLLIN | POLLOUT | POLLHUP | POLLERR))) continue;
      Op *o = &g_ops[idx[j]];
      if (o->hleft > 0) {               /* the size header first */
        char *h = (char *)&o->hdr + (8 - o->hleft);
        ssize_t k = o->send ? write(pf[j].fd, h, o->hleft) : read(pf[j].fd, h, o->hleft);
        if (k < 0) { if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) continue; rc = 7; break; }
        if (k == 0 && !o->send) { rc = 8; break; }
        o->hleft -= (int)k;
        if (o->hleft == 0 && !o->send && o->hdr != o->size) {
          fprintf(stderr, "nccl_emu: rank %d expected %zu bytes from rank %d, the sender posted %zu\n",
                  o->c->rank, o->size, o->peer, o->hdr);
          rc = 14; break;
        }
        continue;
      }
      ssize_t k = o->send ? write(pf[j].fd, o->p, o->left) : read(pf[j].fd, o->p, o->left);
      if (k < 0) { if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) continue; rc = 7; break; }
      if (k == 0 && !o->send) { rc = 8; break; }
      o->p += k; o->left -= (size_t)k;
      if (!o->left) --live;
    }
  }
  for (int i = 0; i < g_nops; ++i) {
    int fl = fcntl(g_ops[i].c->fd[g_ops[i].peer], F_GETFL, 0);
    fcntl(g_ops[i].c->fd[g_ops[i].peer], F_SETFL, fl & ~O_NONBLOCK);
  }
  g_nops = 0;
  return rc;
}

int ncclGroupStart(void) { ++g_depth; return 0; }
int ncclGroupEnd(void) {
  if (--g_depth > 0) return 0;
  return progress_all();
}

static size_t tsize(int dtype) { return dtype == 8 ? 8 : (dtype == 7 ? 4 : 1); }

static int post(int send, void *p, size_t count, int dtype, int peer, Comm *c) {
  if (peer == c->rank || peer < 0 || peer >= c->nranks || g_nops >= 4096) return 9;
  if (count == 0) return 0;
  g_ops[g_nops].send = send; g_ops[g_nops].p = (char *)p; g_ops[g_nops].left = count*tsize(dtype);
  g_ops[g_nops].peer = peer; g_ops[g_nops].c = c;
  g_ops[g_nops].size = g_ops[g_nops].left; g_ops[g_nops].hdr = g_ops[g_nops].left; g_ops[g_nops].hleft = 8;
  ++g_nops;
  return g_depth > 0 ? 0 : progress_all();
}
int ncclSend(const void *p, size_t count, int dtype, int peer, ncclComm_t c, void *stream) {
  (void)stream; return post(1, (void *)p, count, dtype, peer, c);
}
int ncclRecv(void *p, size_t count, int dtype, int peer, ncclComm_t c, void *stream) {
  (void)stream; return post(0, p, count, dtype, peer, c);
}

int ncclAllReduce(const void *in, void *out, size_t count, int dtype, int op, ncclComm_t c,
                  void *stream) {
  (void)stream;
  if (dtype != 8 || (op != 0 && op != 3) || g_depth > 0) return 10;
  double *o = (double *)out;
  if (out != in) memmove(out, in, count*8);
  if (c->nranks == 1) return 0;
  if (c->rank == 0) {
    double *tmp = (double *)malloc(count*8 ? count*8 : 8);
    for (int r = 1; r < c->nranks; ++r) {
      if (xread(c->fd[r], tmp, count*8)) { free(tmp); return 11; }
      for (size_t i = 0; i < count; ++i) {
        if (op == 0) o[i] += tmp[i];
        else if (tmp[i] < o[i]) o[i] = tmp[i];
      }
    }
    free(tmp);
    for (int r = 1; r < c->nranks; ++r) if (xwrite(c->fd[r], o, count*8)) return 12;
  } else {
    if (xwrite(c->fd[0], o, count*8) || xread(c->fd[0], o, count*8)) return 13;
  }
  return 0;
}
