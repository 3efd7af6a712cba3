/* include/bridge.h — the drop-in boundary.
 *
 * The nine extern "C" entry points Booster's Go server binds through cgo. Prototypes are
 * identical to the reference's cpp/bridge.h:132-165 (definitions cpp/bridge.cpp:697-835) and to the
 * cgo preamble the caller re-declares them with (pkg/server/server.go:7-36, pkg/booster/booster.go:15-22).
 * libbooster_b200.so exports exactly these symbols, so the Go binary links against it instead of the
 * prebuilt cpp objects listed in booster.go:8 LDFLAGS and is otherwise unchanged.
 *
 * Plain C: no C++ or torch types cross this boundary.
 */
#ifndef BOOSTER_B200_BRIDGE_H
#define BOOSTER_B200_BRIDGE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cpp/bridge.h:134, cpp/bridge.cpp:711-720 — once per pod; `swap` (session dir) is accepted and ignored
 * exactly as the reference does (its session code is commented out); `debug` containing "cuda" un-mutes logs.
 * NB the Go preamble declares the return type as `void *` and ignores it (server.go:10,553). */
void init(char * swap, char * debug);

/* cpp/bridge.h:136-148, cpp/bridge.cpp:723-786 — load the GGUF model and create the decode context for pod
 * idx in [0,8). gpu1..gpu4 are layer-split proportions over devices 0..3 (server.go:514-530); their sum is
 * n_gpu_layers (bridge.cpp:746-750). Returns an opaque context or NULL. The pointer lives for the process
 * lifetime (the ABI has no free). More than 4 GPUs: env BOOSTER_B200_SPLIT="p0,p1,...,p7" (additive). */
void * initContext(
    int idx,
    char * modelName,
    int threads,
    int batch_size,
    int gpu1, int gpu2, int gpu3, int gpu4,
    int context, int predict,
    int32_t mirostat, float mirostat_tau, float mirostat_eta,
    float temperature, int top_k, float top_p,
    float typical_p,
    float repetition_penalty, int penalty_last_n,
    int32_t janus, int32_t depth, float scale, float hi, float lo,
    uint32_t seed,
    char * debug);

/* cpp/bridge.h:150-155, cpp/bridge.cpp:788-800 → do_inference :175-658 — blocking; runs the whole generation.
 * Returns n_p_eval + n_eval (bridge.cpp:657); 0 when the prompt does not fit n_ctx-4 (:382-386);
 * 1 when a decode step failed (:556-558). All char* are copied, never retained. */
int64_t doInference(
    int idx,
    void * ctx,
    char * jobID,
    char * sessionID,
    char * prompt);

/* cpp/bridge.h:157, cpp/bridge.cpp:802-804 — asynchronous stop request for pod idx (any thread). */
void stopInference(int idx);

/* cpp/bridge.h:158, cpp/bridge.cpp:807-811 → :662-667 — full text so far (prompt + generated pieces);
 * callable concurrently with doInference; the returned pointer stays valid until the next status() call
 * for the same job (Go copies immediately, server.go:842). */
const char * status(char * jobID);

/* cpp/bridge.h:159-161, cpp/bridge.cpp:813-828 — integer milliseconds per token (prompt / generation)
 * and the prompt token count, as the reference stores them (bridge.cpp:650-655). */
int64_t promptEval(char * jobID);
int64_t getPromptTokenCount(char * jobID);
int64_t timing(char * jobID);

/* cpp/bridge.h:162, cpp/bridge.cpp:830-834 — the seed used for the job (bridge.cpp:216-221). */
uint32_t getSeed(char * jobID);

#ifdef __cplusplus
}
#endif
#endif /* BOOSTER_B200_BRIDGE_H */
