/*
 * mpi_shim.c -- fork()/socketpair() mini-MPI.  TEST INFRASTRUCTURE ONLY.
 *
 * Implements the 17 MPI calls used by the reference's short-range path
 * (/root/reference/src/{domains,fmm,remotes,toptree,initial}.c; see SURVEY.md
 * Appendix C) so that the unmodified reference can be executed at NP >= 1 on a
 * host with no MPI installation.  NP comes from the environment variable
 * PN_SHIM_NP (default 1).  MPI_Init() forks NP-1 children; every process then
 * returns from MPI_Init() as one rank of a full socketpair mesh.
 *
 * Semantics: sends are buffered (MPI_Isend copies the payload), matching is
 * (source, tag) in arrival order, all progress is made inside the shim calls.
 * Datatypes are their byte size (see mpi.h).
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#define MAXNP 128
#define TAG_INTERNAL (1 << 20)

typedef struct Msg {
    struct Msg *next;
    int tag;
    long nbytes;
    char *data;
} Msg;

typedef struct OutBuf {
    struct OutBuf *next;
    long nbytes, off;
    char *data;
} OutBuf;

typedef struct {
    int tag;
    int pad;
    long nbytes;
} Hdr;

static int g_np = 1, g_rank = 0, g_init = 0;
static int g_fd[MAXNP];
static pid_t g_child[MAXNP];
static Msg *g_inq_head[MAXNP], *g_inq_tail[MAXNP];
static OutBuf *g_out_head[MAXNP], *g_out_tail[MAXNP];
/* partial-receive state per peer */
static Hdr g_rh[MAXNP];
static long g_rh_got[MAXNP];
static char *g_rp[MAXNP];
static long g_rp_got[MAXNP];

long pn_shim_recv_seq[256];
long pn_shim_recv_bytes[256];

int pn_shim_world_size(void) { return g_np; }
int pn_shim_world_rank(void) { return g_rank; }

static void die(const char *what) {
    fprintf(stderr, "[mpi_shim rank %d] fatal: %s (errno=%d %s)\n", g_rank, what, errno, strerror(errno));
    _exit(97);
}

static void inq_push(int src, int tag, long nbytes, char *data) {
    Msg *m = (Msg *)malloc(sizeof(Msg));
    m->next = NULL; m->tag = tag; m->nbytes = nbytes; m->data = data;
    if (g_inq_tail[src]) g_inq_tail[src]->next = m; else g_inq_head[src] = m;
    g_inq_tail[src] = m;
}

static Msg *inq_pop(int src, int tag) {
    Msg *prev = NULL, *m = g_inq_head[src];
    while (m) {
        if (m->tag == tag) {
            if (prev) prev->next = m->next; else g_inq_head[src] = m->next;
            if (g_inq_tail[src] == m) g_inq_tail[src] = prev;
            return m;
        }
        prev = m; m = m->next;
    }
    return NULL;
}

static void do_read(int p) {
    for (;;) {
        if (g_rh_got[p] < (long)sizeof(Hdr)) {
            ssize_t k = read(g_fd[p], (char *)&g_rh[p] + g_rh_got[p], sizeof(Hdr) - g_rh_got[p]);
            if (k < 0) { if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) return; die("read hdr"); }
            if (k == 0) return; /* peer closed */
            g_rh_got[p] += k;
            if (g_rh_got[p] < (long)sizeof(Hdr)) return;
            g_rp[p] = (char *)malloc(g_rh[p].nbytes > 0 ? g_rh[p].nbytes : 1);
            g_rp_got[p] = 0;
        }
        if (g_rp_got[p] < g_rh[p].nbytes) {
            ssize_t k = read(g_fd[p], g_rp[p] + g_rp_got[p], g_rh[p].nbytes - g_rp_got[p]);
            if (k < 0) { if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) return; die("read payload"); }
            if (k == 0) return;
            g_rp_got[p] += k;
            if (g_rp_got[p] < g_rh[p].nbytes) return;
        }
        inq_push(p, g_rh[p].tag, g_rh[p].nbytes, g_rp[p]);
        g_rp[p] = NULL; g_rh_got[p] = 0; g_rp_got[p] = 0;
    }
}

static void do_write(int p) {
    while (g_out_head[p]) {
        OutBuf *b = g_out_head[p];
        ssize_t k = write(g_fd[p], b->data + b->off, b->nbytes - b->off);
        if (k < 0) { if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) return; die("write"); }
        b->off += k;
        if (b->off < b->nbytes) return;
        g_out_head[p] = b->next;
        if (!g_out_head[p]) g_out_tail[p] = NULL;
        free(b->data); free(b);
    }
}

static void progress(int block) {
    struct pollfd pf[MAXNP];
    int idx[MAXNP], n = 0, p;
    for (p = 0; p < g_np; p++) {
        if (p == g_rank) continue;
        pf[n].fd = g_fd[p];
        pf[n].events = POLLIN | (g_out_head[p] ? POLLOUT : 0);
        pf[n].revents = 0;
        idx[n++] = p;
    }
    if (n == 0) return;
    int r = poll(pf, n, block ? 1000 : 0);
    if (r < 0) { if (errno == EINTR) return; die("poll"); }
    for (int i = 0; i < n; i++) {
        if (pf[i].revents & POLLIN) do_read(idx[i]);
        if (pf[i].revents & POLLOUT) do_write(idx[i]);
        if ((pf[i].revents & (POLLHUP | POLLERR)) && !(pf[i].revents & POLLIN)) {
            fprintf(stderr, "[mpi_shim rank %d] peer %d hung up\n", g_rank, idx[i]);
            _exit(98);
        }
    }
}

static void send_bytes(const void *buf, long nbytes, int dest, int tag) {
    if (dest == g_rank) {
        char *d = (char *)malloc(nbytes > 0 ? nbytes : 1);
        memcpy(d, buf, nbytes);
        inq_push(dest, tag, nbytes, d);
        return;
    }
    OutBuf *b = (OutBuf *)malloc(sizeof(OutBuf));
    b->next = NULL; b->off = 0; b->nbytes = sizeof(Hdr) + nbytes;
    b->data = (char *)malloc(b->nbytes);
    Hdr h; h.tag = tag; h.pad = 0; h.nbytes = nbytes;
    memcpy(b->data, &h, sizeof(Hdr));
    memcpy(b->data + sizeof(Hdr), buf, nbytes);
    if (g_out_tail[dest]) g_out_tail[dest]->next = b; else g_out_head[dest] = b;
    g_out_tail[dest] = b;
    do_write(dest);
}

static long recv_bytes(void *buf, long maxbytes, int src, int tag) {
    Msg *m;
    while (!(m = inq_pop(src, tag))) progress(1);
    long n = m->nbytes < maxbytes ? m->nbytes : maxbytes;
    memcpy(buf, m->data, n);
    long got = m->nbytes;
    free(m->data); free(m);
    return got;
}

static void flush_to(int dest) {
    if (dest == g_rank) return;
    while (g_out_head[dest]) progress(1);
}

int MPI_Init(int *argc, char ***argv) {
    (void)argc; (void)argv;
    if (g_init) return 0;
    g_init = 1;
    const char *e = getenv("PN_SHIM_NP");
    g_np = e ? atoi(e) : 1;
    if (g_np < 1) g_np = 1;
    if (g_np > MAXNP) { fprintf(stderr, "mpi_shim: NP too large\n"); exit(2); }
    g_rank = 0;
    if (g_np == 1) return 0;
    signal(SIGPIPE, SIG_IGN);
    static int pairs[MAXNP][MAXNP][2];
    for (int i = 0; i < g_np; i++)
        for (int j = i + 1; j < g_np; j++) {
            if (socketpair(AF_UNIX, SOCK_STREAM, 0, pairs[i][j]) < 0) die("socketpair");
            int big = 4 << 20;
            setsockopt(pairs[i][j][0], SOL_SOCKET, SO_SNDBUF, &big, sizeof big);
            setsockopt(pairs[i][j][1], SOL_SOCKET, SO_SNDBUF, &big, sizeof big);
        }
    fflush(stdout); fflush(stderr);
    for (int r = 1; r < g_np; r++) {
        pid_t pid = fork();
        if (pid < 0) die("fork");
        if (pid == 0) { g_rank = r; break; }
        g_child[r] = pid;
    }
    for (int i = 0; i < g_np; i++)
        for (int j = i + 1; j < g_np; j++) {
            if (i == g_rank) { g_fd[j] = pairs[i][j][0]; close(pairs[i][j][1]); }
            else if (j == g_rank) { g_fd[i] = pairs[i][j][1]; close(pairs[i][j][0]); }
            else { close(pairs[i][j][0]); close(pairs[i][j][1]); }
        }
    for (int p = 0; p < g_np; p++) {
        if (p == g_rank) continue;
        int fl = fcntl(g_fd[p], F_GETFL, 0);
        fcntl(g_fd[p], F_SETFL, fl | O_NONBLOCK);
    }
    return 0;
}

int MPI_Finalize(void) {
    if (g_np == 1) return 0;
    MPI_Barrier(MPI_COMM_WORLD);
    for (int p = 0; p < g_np; p++) flush_to(p);
    if (g_rank == 0) {
        for (int r = 1; r < g_np; r++) { int st; waitpid(g_child[r], &st, 0); }
    } else {
        fflush(stdout); fflush(stderr);
        _exit(0);
    }
    return 0;
}

int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", code); _exit(code ? code : 1); return 0; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = g_np; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = g_rank; return 0; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) { *out = c; return 0; }
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newtype) { *newtype = count * old; return 0; }
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }

int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *req) {
    (void)c;
    send_bytes(buf, (long)count * t, dest, tag);
    if (req) *req = dest;
    return 0;
}

int MPI_Wait(MPI_Request *req, MPI_Status *st) {
    (void)st;
    if (req && *req >= 0 && *req < g_np) flush_to(*req);
    return 0;
}

int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) {
    (void)c;
    long got = recv_bytes(buf, (long)count * t, src, tag);
    if (tag >= 0 && tag < 256) { pn_shim_recv_seq[tag]++; pn_shim_recv_bytes[tag] = got; }
    if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
    return 0;
}

int MPI_Barrier(MPI_Comm c) {
    (void)c;
    if (g_np == 1) return 0;
    char tok = 0;
    if (g_rank == 0) {
        for (int p = 1; p < g_np; p++) recv_bytes(&tok, 1, p, TAG_INTERNAL + 1);
        for (int p = 1; p < g_np; p++) send_bytes(&tok, 1, p, TAG_INTERNAL + 2);
        for (int p = 1; p < g_np; p++) flush_to(p);
    } else {
        send_bytes(&tok, 1, 0, TAG_INTERNAL + 1);
        flush_to(0);
        recv_bytes(&tok, 1, 0, TAG_INTERNAL + 2);
    }
    return 0;
}

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, MPI_Comm c) {
    (void)c;
    long sb = (long)scount * st, rb = (long)rcount * rt;
    for (int p = 0; p < g_np; p++) send_bytes(sbuf, sb, p, TAG_INTERNAL + 3);
    for (int p = 0; p < g_np; p++) recv_bytes((char *)rbuf + p * rb, rb, p, TAG_INTERNAL + 3);
    for (int p = 0; p < g_np; p++) flush_to(p);
    return 0;
}

int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, MPI_Comm c) {
    (void)c;
    long sb = (long)scount * st, rb = (long)rcount * rt;
    for (int p = 0; p < g_np; p++) send_bytes((const char *)sbuf + p * sb, sb, p, TAG_INTERNAL + 4);
    for (int p = 0; p < g_np; p++) recv_bytes((char *)rbuf + p * rb, rb, p, TAG_INTERNAL + 4);
    for (int p = 0; p < g_np; p++) flush_to(p);
    return 0;
}

int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype st,
                  void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rt, MPI_Comm c) {
    (void)c;
    for (int p = 0; p < g_np; p++)
        send_bytes((const char *)sbuf + (long)sdispls[p] * st, (long)scounts[p] * st, p, TAG_INTERNAL + 5);
    for (int p = 0; p < g_np; p++)
        recv_bytes((char *)rbuf + (long)rdispls[p] * rt, (long)rcounts[p] * rt, p, TAG_INTERNAL + 5);
    for (int p = 0; p < g_np; p++) flush_to(p);
    return 0;
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    /* the reference reduces doubles only (src/photoNs.c:277-281); ints are supported for the harness */
    long nb = (long)count * t;
    char *all = (char *)malloc(nb * g_np);
    MPI_Allgather(sbuf, count, t, all, count, t, c);
    for (int i = 0; i < count; i++) {
        if (t == 8) {
            double acc = ((double *)all)[i];
            for (int p = 1; p < g_np; p++) {
                double v = ((double *)(all + p * nb))[i];
                if (op == MPI_SUM) acc += v; else if (op == MPI_MAX) acc = v > acc ? v : acc; else acc = v < acc ? v : acc;
            }
            ((double *)rbuf)[i] = acc;
        } else {
            int acc = ((int *)all)[i];
            for (int p = 1; p < g_np; p++) {
                int v = ((int *)(all + p * nb))[i];
                if (op == MPI_SUM) acc += v; else if (op == MPI_MAX) acc = v > acc ? v : acc; else acc = v < acc ? v : acc;
            }
            ((int *)rbuf)[i] = acc;
        }
    }
    free(all);
    return 0;
}
