/* libctcb200 — C ABI of the B200-native CTC acoustic hot path.
 *
 * Every entry point takes raw device pointers, sizes and the CUDA stream to enqueue on, returns 0 on
 * success or a negative ctcb200 status, never throws and never allocates memory the caller must free.
 * The message for the last failure on the calling thread is available from ctcb200_last_error().
 * All work is asynchronous on `stream`; nothing here synchronises the device.
 *
 * The reference (Diamondfan/CTC_pytorch) is pure Python on top of PyTorch, so it has no FFI of its own;
 * each function below names the reference call site (file:line under /root/reference) whose library call
 * it replaces. The Python mirror of the reference's classes (ctc_pytorch_b200/*.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 */
#ifndef CTCB200_H_
#define CTCB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(CTCB200_BUILD)
#define CTCB200_API __attribute__((visibility("default")))
#else
#define CTCB200_API
#endif

/* CUDA stream handle (cudaStream_t) passed as an opaque pointer so this header needs no CUDA include. */
typedef void* ctcb200_stream_t;

enum {
    CTCB200_OK = 0,
    CTCB200_ERR_INVALID = -1,
    CTCB200_ERR_CUDA = -2,
    CTCB200_ERR_DRIVER = -3,
    CTCB200_ERR_TIMEOUT = -4
};

CTCB200_API const char* ctcb200_last_error(void);
CTCB200_API int ctcb200_version(void);

/* ---- CTC loss: replaces nn.CTCLoss(reduction='sum') fwd/bwd, timit/steps/train_ctc.py:144,47,63 ------
 * log_probs [T,N,C] f32; targets [N,*] int64 zero-padded rows of pitch target_stride
 * (timit/utils/data_loader.py:125,140); input_lengths / target_lengths [N] int64.
 * alpha_ws: caller-provided workspace of ctcb200_ctc_workspace_floats(T,N,max_target_len) floats, kept
 * between fwd and bwd. nll [N] f32 = per-utterance negative log likelihood (+inf if infeasible).
 * bwd writes grad [T,N,C] f32 = grad_scale * grad_nll[n] * (exp(lp) - exp(lcab + nll - lp)), zero for
 * t >= input_length (torch's native convention); grad_nll may be NULL (= all ones). */
CTCB200_API int64_t ctcb200_ctc_workspace_floats(int T, int N, int max_target_len);
CTCB200_API int ctcb200_ctc_loss_fwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                                     const int64_t* input_lengths, const int64_t* target_lengths, int T, int N,
                                     int C, int max_target_len, int blank, float* alpha_ws, float* nll,
                                     ctcb200_stream_t stream);
CTCB200_API int ctcb200_ctc_loss_bwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                                     const int64_t* input_lengths, const int64_t* target_lengths, int T, int N,
                                     int C, int max_target_len, int blank, const float* alpha_ws, const float* nll,
                                     const float* grad_nll, float grad_scale, float* grad, ctcb200_stream_t stream);

/* ---- greedy path: replaces torch.max(out,-1) (train_ctc.py:51), the collapse loop of
 * CTC_Model.compute_wer (timit/models/model_ctc.py:187-202) and GreedyDecoder.decode
 * (timit/utils/ctcDecoder.py:162-166). idx_nt [N,T] int32 frame arg-max (first index on ties);
 * labels_nt [N,T] int32 collapsed labels, label_lengths [N] int32. */
CTCB200_API int ctcb200_argmax(const float* log_probs, int T, int N, int C, int* idx_nt, float* max_nt,
                               ctcb200_stream_t stream);
CTCB200_API int ctcb200_greedy_decode(const float* log_probs, const int64_t* lengths, int T, int N, int C, int blank,
                                      int* idx_nt, int* labels_nt, int* label_lengths, ctcb200_stream_t stream);

/* ---- dense GEMM on tcgen05: C[M,N] (+)= A[M,K] * B[N,K]^T, A/B bf16 with K contiguous (pitches lda/ldb in
 * elements, multiples of 8), fp32 accumulate, C f32 (out_bf16=0) or bf16 (1) with pitch ldc.
 * a_koff/b_koff shift the K window of each operand (used for the h_{t-1} shift of dW_hh).
 * tile_n: 0 = auto, else 64/128/256. Carries the contractions behind nn.LSTM / nn.Linear at
 * timit/models/model_ctc.py:23-26,33,136-139. */
CTCB200_API int ctcb200_gemm_tn_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                                     int M, int N, int K, int a_koff, int b_koff, int out_bf16, int accumulate,
                                     int tile_n, ctcb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCB200_H_ */
