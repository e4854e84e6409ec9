/* batcher_stress.c -- the group-commit protocol of zg_check_bulk / zg_lookup_resources (csrc/capi.cu: queue, leader,
 * hand-over to the head of the queue, one wake-up per finished caller) under many threads WITHOUT a GPU: a host-only
 * engine runs the whole protocol and answers every group with ZG_ECUDA (the library has no CPU evaluation path). A lost
 * wake-up or a broken hand-over shows as a hang (the test runs under a timeout), a mixed-up group as a wrong return
 * code or another thread's error text. Test infrastructure: not part of libzgpu.so. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "zgpu.h"

static zg_engine *E;
static int THREADS = 64;
static double SECONDS = 2.0;
static volatile int workers_stop;
static long total_calls, total_writes;
static int TIGHT_WRITER;
static pthread_barrier_t start;
static const uint64_t BIG_N = 1ull << 23;
static zg_check *BIG_ITEMS;
static uint8_t *BIG_OUT;

static void *worker(void *arg) {
  const long id = (long)arg;
  zg_check items[8];
  uint8_t out[8];
  memset(items, 0, sizeof items);
  pthread_barrier_wait(&start);
  int i = 0;
  for (; !workers_stop; ++i) {
    int rc;
    if ((i + id) % 3 == 0) {
      uint32_t ids[4];
      uint64_t n = 0;
      rc = zg_lookup_resources(E, 1, 0, 0, (uint32_t)i, ZG_SREL_NONE, ids, 4, &n);
    } else if ((i + id) % 4 == 1) {
      /* half a launch's worth of items (never read: the group fails before any copy): two such requests fill a
       * group, so a queue of them is answered group after group, each leader handing over to the next */
      rc = zg_check_bulk(E, BIG_ITEMS, BIG_N, BIG_OUT);
    } else {
      rc = zg_check_bulk(E, items, 1 + (uint64_t)((i + id) % 8), out);
    }
    if ((i & 63) == 0) {  /* let the writer in: its publishes are what makes callers queue */
      struct timespec ts = {0, 200000};
      nanosleep(&ts, NULL);
    }
    if (rc != ZG_ECUDA) {
      printf("thread %ld call %d: rc %d (%s)\n", id, i, rc, zg_last_error());
      return (void *)1;
    }
    if (!strstr(zg_last_error(), "host-only")) {
      printf("thread %ld call %d: error text `%s`\n", id, i, zg_last_error());
      return (void *)1;
    }
    if ((i & 1023) == 0) {  /* an unrelated failing call in between: its text must not leak into the next check */
      if (zg_check_bulk(E, NULL, 3, out) != ZG_EINVAL) return (void *)1;
    }
  }
  __atomic_fetch_add(&total_calls, i, __ATOMIC_RELAXED);
  return NULL;
}

/* Holds the device lock for milliseconds at a time (a write publishes under it), so that check and lookup leaders wait
 * for the device with callers queueing behind them: the hand-over and the per-caller wake-ups are what runs then. */
static volatile int writer_stop;
static void *writer(void *arg) {
  (void)arg;
  char id[32];
  long n = 0;
  while (!writer_stop) {
    zg_update_str up[64];
    char ids[64][32];
    for (int k = 0; k < 64; ++k) {
      snprintf(ids[k], sizeof ids[k], "d%ld", (n * 64 + k) % 20000);
      zg_update_str u = {{"doc", ids[k], "viewer", "user", "u1", ""}, 0, ZG_OP_TOUCH};
      up[k] = u;
    }
    if (zg_write_relationships(E, up, 64, NULL, 0)) {
      printf("writer: %s\n", zg_last_error());
      return (void *)1;
    }
    ++n;
    if (!TIGHT_WRITER) { /* paced like RPCs; argv[3] = "tight": back-to-back writes, which starved every other caller
                            for seconds before the device lock became FIFO */
      struct timespec ts = {0, 2000000};
      nanosleep(&ts, NULL);
    }
    __atomic_fetch_add(&total_writes, 1, __ATOMIC_RELAXED);
  }
  (void)id;
  return NULL;
}

int main(int argc, char **argv) {
  if (argc > 1) THREADS = atoi(argv[1]);
  if (argc > 2) SECONDS = atof(argv[2]);
  TIGHT_WRITER = argc > 3 && !strcmp(argv[3], "tight");
  zg_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.flags = ZG_FLAG_HOST_ONLY;
  if (zg_engine_create(&cfg, &E)) return printf("create: %s\n", zg_last_error()), 1;
  const char *schema = "definition user {}\ndefinition doc { relation viewer: user  permission view = viewer }";
  if (zg_load_schema(E, schema, strlen(schema))) return printf("schema: %s\n", zg_last_error()), 1;
  {  /* a store large enough that every publish (a host CSR build here) holds the device lock for milliseconds */
    const uint32_t N = 300000;
    zg_tuple *t = calloc(N, sizeof *t);
    const int doc = zg_type_id(E, "doc"), user = zg_type_id(E, "user");
    const uint32_t u1 = zg_intern_object(E, user, "u1");
    char name[32];
    for (uint32_t i = 0; i < N; ++i) {
      snprintf(name, sizeof name, "d%u", i);
      t[i].res = zg_intern_object(E, doc, name);
      t[i].subj = u1;
      t[i].rel = (uint16_t)zg_slot_id(E, doc, "viewer");
      t[i].stype = (uint16_t)user;
      t[i].srel = ZG_SREL_NONE;
    }
    if (zg_load_tuples(E, t, NULL, N) || zg_publish(E)) return printf("load: %s\n", zg_last_error()), 1;
    free(t);
  }
  BIG_ITEMS = calloc(BIG_N, sizeof(zg_check)); /* untouched pages cost nothing */
  BIG_OUT = calloc(BIG_N, 1);
  if (!BIG_ITEMS || !BIG_OUT) return printf("no memory\n"), 1;
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)THREADS), wr;
  pthread_barrier_init(&start, NULL, (unsigned)THREADS);
  pthread_create(&wr, NULL, writer, NULL);
  for (long i = 0; i < THREADS; ++i) pthread_create(&th[i], NULL, worker, (void *)i);
  {
    struct timespec ts = {(time_t)SECONDS, (long)((SECONDS - (time_t)SECONDS) * 1e9)};
    nanosleep(&ts, NULL);
    workers_stop = 1;
    writer_stop = 1;
  }
  int bad = 0;
  for (int i = 0; i < THREADS; ++i) {
    void *r;
    pthread_join(th[i], &r);
    bad += r != NULL;
  }
  writer_stop = 1;
  {
    void *r;
    pthread_join(wr, &r);
    bad += r != NULL;
  }
  zg_engine_destroy(E);
  if (bad) return printf("%d threads failed\n", bad), 1;
  printf("batcher stress ok: %d threads, %ld calls and %ld writes in %.1f s\n", THREADS, total_calls, total_writes, SECONDS);
  return 0;
}
