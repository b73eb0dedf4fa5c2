/*
 * racformer_tools.h -- C ABI of tools/lib/libracformer_tools.so: measurement aids only. Nothing in racformer_b200/
 * links or loads this library; it is built on demand by tools/tools_lib.py and used by tools/gather_ceiling.py.
 */
#ifndef RACFORMER_TOOLS_H_
#define RACFORMER_TOOLS_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef void* racf_stream_t;
#define RACF_OK                 0
#define RACF_ERR_NULL_POINTER  (-1)
#define RACF_ERR_BAD_SHAPE     (-3)

/*
 * Random 512-byte coalesced row reads (the request shape of one bilinear cell row) over buf[0 : num_rows * 512 B],
 * total_rows reads, `ilp` independent loads in flight per warp (1,2,4,8,16): the achievable gather bandwidth for
 * HBM- and L2-sized footprints.
 */
int racf_bench_gather_ceiling(const float* buf, long long num_rows, long long total_rows, int ilp,
                              float* sink, racf_stream_t stream);

/* The same for the scatter -- red.global.add.v4.f32 on random 512-byte rows of buf. */
int racf_bench_scatter_ceiling(float* buf, long long num_rows, long long total_rows, racf_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif
