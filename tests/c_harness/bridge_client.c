/* tests/c_harness/bridge_client.c — the cgo caller's view of libbooster_b200.so, in plain C.
 *
 * The Go toolchain is absent from the build image, so this stands in for the cgo preamble of pkg/server/server.go:7-36 /
 * pkg/booster/booster.go:15-22: it includes include/bridge.h (and the additive include/booster_b200.h) from a C translation
 * unit — proving both headers are plain C — links against the shared library the way `#cgo LDFLAGS: -lbooster_b200` would,
 * and drives the nine symbols in the order server.go does: init -> initContext -> doInference with a concurrent status()
 * poller thread -> promptEval / getPromptTokenCount / timing / getSeed.
 *
 *   bridge_client                    no model: the calls a server makes on a bad configuration (NULL / 0, never a crash)
 *   bridge_client MODEL.gguf PROMPT  a real job on GPU 0 (needs a CUDA device: the library has no CPU path)
 */
#include <pthread.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include "bridge.h"
#include "booster_b200.h"

static volatile int g_done = 0;
static char g_job[] = "c-harness-job";

static void * poller(void * arg) {
    (void) arg;
    size_t seen = 0;
    while (!g_done) {
        const char * s = status(g_job);             /* what server.go:842 copies with C.GoString */
        const size_t n = strlen(s);
        if (n > seen) seen = n;
        usleep(200);
    }
    return (void *) seen;
}

int main(int argc, char ** argv) {
    char empty[] = "", swap[] = "/tmp", none[] = "/nonexistent/model.gguf", nojob[] = "no-such-job";
    init(swap, empty);
    printf("version: %s, devices: %d\n", b200_version(), b200_device_count());
    if (argc < 3) {
        /* a pod whose model cannot be loaded: NULL, and every query on an unknown job answers with the reference's zero values */
        void * ctx = initContext(0, none, 1, 0, 100, 0, 0, 0, 64, 8, 0, 0.0f, 0.0f, 0.0f, 1, 1.0f, 1.0f, 1.0f, 0, 1, 200, 1.0f, 1.0f, 1.0f, 42u, empty);
        if (ctx) { printf("FAIL: initContext returned a context for a missing model\n"); return 1; }
        if (doInference(0, ctx, g_job, empty, empty) != 0) { printf("FAIL: doInference on a NULL context\n"); return 1; }
        if (status(nojob)[0] != 0 || promptEval(nojob) != 0 || getPromptTokenCount(nojob) != 0 || timing(nojob) != 0 || getSeed(nojob) != 0) {
            printf("FAIL: unknown job\n"); return 1;
        }
        stopInference(0); stopInference(-1); stopInference(99);
        printf("ok: bad configuration handled\n");
        return 0;
    }
    void * ctx = initContext(0, argv[1], 1, 0, 100, 0, 0, 0, 256, 16, 0, 0.0f, 0.0f, 0.0f, 1, 1.0f, 1.0f, 1.0f, 0, 1, 200, 1.0f, 1.0f, 1.0f, 42u, empty);
    if (!ctx) { printf("FAIL: initContext(%s)\n", argv[1]); return 1; }
    pthread_t th;
    pthread_create(&th, NULL, poller, NULL);
    const long long n = (long long) doInference(0, ctx, g_job, empty, argv[2]);
    g_done = 1;
    void * seen = NULL;
    pthread_join(th, &seen);
    printf("doInference = %lld, prompt tokens = %lld, seed = %u, text = \"%s\" (poller saw %zu bytes)\n", n,
           (long long) getPromptTokenCount(g_job), getSeed(g_job), status(g_job), (size_t) seen);
    printf("promptEval = %lld ms/token, timing = %lld ms/token\n", (long long) promptEval(g_job), (long long) timing(g_job));
    return n > 1 ? 0 : 1;
}
