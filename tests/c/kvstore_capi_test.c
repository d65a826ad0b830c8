/* C program against the handle-based store API of libdlcfn_rt.so (no Python): local store with an updater, and - when the
 * DMLC_* environment of a launcher is present - a dist_sync worker.  Built and run by tests/test_kvstore.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern const char* DLCFNKVGetLastError(void);
extern int DLCFNKVStoreCreate(const char* type, void** out);
extern int DLCFNKVStoreFree(void* h);
extern int DLCFNKVStoreGetType(void* h, const char** out);
extern int DLCFNKVStoreGetRank(void* h, int* out);
extern int DLCFNKVStoreGetGroupSize(void* h, int* out);
extern int DLCFNKVStoreBarrier(void* h);
extern int DLCFNKVStoreIsWorkerNode(int* out);
extern int DLCFNKVStoreRunServer(void);
extern int DLCFNKVStoreInit(void* h, int num, const int* keys, const void* const* vals, const size_t* lens, const int* dtypes);
extern int DLCFNKVStorePush(void* h, int num, const int* keys, const void* const* vals, int priority);
extern int DLCFNKVStorePull(void* h, int num, const int* keys, void* const* outs, int priority);
typedef void (*KVUpdater)(int key, const void* recv, void* local, size_t n, int dtype, void* user);
extern int DLCFNKVStoreSetUpdater(void* h, KVUpdater f, void* user);

#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "FAILED %s: %s\n", #x, DLCFNKVGetLastError()); return 1; } } while (0)

static void sgd(int key, const void* recv, void* local, size_t n, int dtype, void* user) {
  (void)key; (void)dtype;
  const float lr = *(const float*)user;
  const float* g = (const float*)recv;
  float* w = (float*)local;
  for (size_t i = 0; i < n; ++i) w[i] -= lr * g[i];
}

int main(int argc, char** argv) {
  const char* type = argc > 1 ? argv[1] : "local";
  int worker = 1;
  DLCFNKVStoreIsWorkerNode(&worker);
  if (!worker) {                    /* the launcher starts the same binary for the scheduler and the servers */
    CHECK(DLCFNKVStoreRunServer());
    return 0;
  }
  void* kv = NULL;
  CHECK(DLCFNKVStoreCreate(type, &kv));
  const char* t; int rank, size;
  CHECK(DLCFNKVStoreGetType(kv, &t));
  CHECK(DLCFNKVStoreGetRank(kv, &rank));
  CHECK(DLCFNKVStoreGetGroupSize(kv, &size));
  enum { N = 1000 };
  static float w[N], g1[N], g2[N], out[N];
  for (int i = 0; i < N; ++i) { w[i] = 1.0f; g1[i] = 1.0f; g2[i] = 2.0f; }
  int key = 3, dtype = 0; size_t len = N;
  const void* vals[1] = {w};
  CHECK(DLCFNKVStoreInit(kv, 1, &key, vals, &len, &dtype));
  if (strncmp(type, "dist", 4) != 0) {
    float lr = 0.5f;
    CHECK(DLCFNKVStoreSetUpdater(kv, sgd, &lr));
    int keys2[2] = {3, 3};
    const void* grads[2] = {g1, g2};                 /* two "devices": summed first, then the updater */
    CHECK(DLCFNKVStorePush(kv, 2, keys2, grads, 0));
    void* outs[1] = {out};
    CHECK(DLCFNKVStorePull(kv, 1, &key, outs, 0));
    for (int i = 0; i < N; ++i) if (out[i] != 1.0f - 0.5f * 3.0f) { fprintf(stderr, "local: out[%d] = %f\n", i, out[i]); return 1; }
  } else {
    /* no optimizer on the servers: a sync round stores the sum over workers (kvstore_dist_server.h:358-361) */
    for (int i = 0; i < N; ++i) g1[i] = (float)(rank + 1);
    const void* grads[1] = {g1};
    CHECK(DLCFNKVStorePush(kv, 1, &key, grads, 0));
    void* outs[1] = {out};
    CHECK(DLCFNKVStorePull(kv, 1, &key, outs, 0));
    float want = size * (size + 1) / 2.0f;
    for (int i = 0; i < N; ++i) if (out[i] != want) { fprintf(stderr, "dist: out[%d] = %f want %f\n", i, out[i], want); return 1; }
    CHECK(DLCFNKVStoreBarrier(kv));
  }
  printf("rank %d/%d %s CAPI_OK\n", rank, size, t);
  fflush(stdout);
  CHECK(DLCFNKVStoreFree(kv));
  return 0;
}
