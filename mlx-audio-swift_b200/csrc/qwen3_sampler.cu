// Test entry for the Qwen3-TTS in-graph sampler kernel (qwen3_sampler.cuh); the product caller is the talker loop in llama.cu.
#include "qwen3_sampler.cuh"

using namespace b2a;

extern "C" {

// Test entry (tests/test_gpu_qwen3_sampler.py): HOST logits [B, V], HOST seen bitmap [B, ceil(V/32)] (nullable, in/out when
// track != 0), tokens_out [B], filtered_out [B, V] (nullable).  One launch of q3s::sample_kernel on device 0.
int32_t b2a_qwen3_sample_test(const float* logits, int32_t B, int32_t V, float temperature, float top_p, int32_t top_k, float min_p,
                              float repetition_penalty, int32_t eos, int32_t suppress_lo, int32_t suppress_hi, uint32_t* seen, int32_t track,
                              uint64_t seed, int32_t step, int32_t* tokens_out, float* filtered_out) {
    return guarded([&] {
        B2A_CHECK(logits && tokens_out && B >= 1 && V >= 1 && V <= q3s::SLOTS, B2A_ERR_INVALID_INPUT, "b2a_qwen3_sample_test: bad argument (vocabulary <= 4096)");
        require_device(0);
        const int words = (V + 31) / 32;
        DBuf<float> dl, df;
        DBuf<unsigned> ds;
        DBuf<int> dt;
        dl.upload(logits, (size_t)B * V);
        dt.alloc(B);
        q3s::Args a{};
        a.logits = dl.p; a.V = V; a.temperature = temperature; a.top_p = top_p; a.top_k = top_k; a.min_p = min_p; a.rep_penalty = repetition_penalty;
        a.eos = eos; a.suppress_lo = suppress_lo; a.suppress_hi = suppress_hi; a.track = track; a.seed = seed; a.step = step; a.tokens = dt.p;
        if (seen) { ds.upload(seen, (size_t)B * words); a.seen = ds.p; }
        if (filtered_out) { df.alloc((size_t)B * V); a.filtered = df.p; }
        B2A_CUDA(cudaDeviceSynchronize());
        q3s::sample_kernel<<<B, q3s::THREADS>>>(a);
        count_launch();
        B2A_CUDA(cudaGetLastError());
        B2A_CUDA(cudaDeviceSynchronize());
        B2A_CUDA(cudaMemcpy(tokens_out, dt.p, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost));
        if (seen) B2A_CUDA(cudaMemcpy(seen, ds.p, (size_t)B * words * sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (filtered_out) B2A_CUDA(cudaMemcpy(filtered_out, df.p, (size_t)B * V * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
