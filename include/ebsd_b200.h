/* libebsd_b200.so -- C ABI of the B200-native EBSD dictionary-indexing hot path.
 *
 * Drop-in boundary for the path DiffractionPatternIndexer.build_dictionary / index_pattern of
 * poyentung/ebsd-vae ("latice").  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference root).  The reference is pure Python; its maintainer binds these
 * with ctypes, exactly as ebsd_vae_b200/_native.py does.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates user-visible memory: outputs and workspaces are caller-owned
 *     (torch tensors in the Python host);
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - return value 0 = OK, negative = error (EBSD_ERR_*); ebsd_last_error() returns a thread-local message;
 *   - sm_100a only: there is no CPU path and no other architecture. On a device that is not
 *     compute capability 10.x every call fails with EBSD_ERR_ARCH.
 */
#ifndef EBSD_B200_H
#define EBSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EBSD_OK 0
#define EBSD_ERR_ARG (-1)
#define EBSD_ERR_CUDA (-2)
#define EBSD_ERR_ARCH (-3)
#define EBSD_ERR_WORKSPACE (-4)

#define EBSD_LATENT_DIM 16
#define EBSD_IMAGE_SIZE 128
#define EBSD_MAX_TOPK 32       /* top_n supported by ebsd_topk / ebsd_consensus */
#define EBSD_N_CONV 10

#define EBSD_PATTERN_U8 0      /* uint8 [B,128,128], value k means k/255 (ToTensor, latice/data_module.py:31) */
#define EBSD_PATTERN_F32 1     /* float32 [B,128,128], used as is (tensor inputs bypass the transform, dp_indexer.py:128-131) */

/* ebsd_quantize_crop source types.  U8 / F32 / F64 alone: ToPILImage semantics (encode_pattern path).  Any type
 * | EBSD_SRC_VIA_F64: the frame is cast to float64 first, as DPdataset.__getitem__ does (latice/data_module.py:132). */
#define EBSD_SRC_U8 0
#define EBSD_SRC_F32 1
#define EBSD_SRC_F64 2
#define EBSD_SRC_I16 3
#define EBSD_SRC_U16 4
#define EBSD_SRC_I32 5
#define EBSD_SRC_I64 6
#define EBSD_SRC_VIA_F64 16

#define EBSD_ANGLE_RADIANS 0   /* Chroma path: threshold compared with radians (latice/index/chroma_db.py:307-310) */
#define EBSD_ANGLE_DEGREES 1   /* FAISS path: np.degrees first (latice/index/faiss_db.py:308-313) */

int ebsd_abi_version(void);
const char *ebsd_last_error(void);
/* Number of kernels this library has launched in this process (monotonic; for bench.py's gpu_launches). */
uint64_t ebsd_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Input transform: 8-bit quantise + centre crop / zero pad to 128 x 128
 *   replaces create_default_transform (latice/data_module.py:17-33) as applied by DPdataset.__getitem__
 *   (data_module.py:122-133) and encode_pattern / encode_patterns_batch (latice/index/dp_indexer.py:124-126,
 *   150-163), up to the uint8 image (ToTensor's k/255 is applied inside the encoder).
 * src: [B,H,W] device array, src_dtype EBSD_SRC_U8 (copied), _F32, _F64 ((x*255).astype(uint8) with numpy/x86
 * semantics: product rounded in the source precision, truncated toward zero, low byte kept; NaN and
 * |x*255| >= 2^31 give 0), or any EBSD_SRC_* | EBSD_SRC_VIA_F64 (cast to float64 first: the dataset path).  Rows [sy, sy+ly) of the source go to rows [dy, dy+ly) of the 128-row output, columns
 * likewise; everything else is zero (torchvision center_crop: see ebsd_vae_b200/transform.py:_axis_window).
 * dst: uint8 [B,128,128], 4-byte aligned.
 * ------------------------------------------------------------------------------------------- */
int ebsd_quantize_crop(const void *src, int src_dtype, int64_t B, int H, int W, int sy, int dy, int ly, int sx, int dx,
                       int lx, uint8_t *dst, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Angle file: the text AFTER the two header lines -> float64 [rows][3] (phi1, Phi, phi2 in degrees)
 *   replaces the regular case of DPDataModule._parse_rotation_angles (latice/data_module.py:87-116: fields separated
 *   by single spaces, empty fields dropped, float(token)).  Host code, no GPU; callable without the Python GIL.
 * Returns the number of rows, or -1 when the text is not "three plain decimal numbers per line" (short / long / blank
 * rows, tabs, nan, underscores, non-ASCII ...) or does not fit `capacity_rows`: the caller then runs its generic
 * parser, which reproduces the reference's NaN padding and error messages (ebsd_vae_b200/transform.py).
 * ------------------------------------------------------------------------------------------- */
int64_t ebsd_parse_angle_text(const char *text, size_t len, double *out, int64_t capacity_rows);

/* ---------------------------------------------------------------------------------------------
 * Encoder: VariationalAutoEncoderRawData.encoder + mu / logvar heads
 *   replaces latice/model.py:55-58 (forward up to logvar), layer plan latice/model.py:93-129,
 *   as called from latice/index/dp_indexer.py:133-137, 177-184, 281-287.
 * Weights are fp32 device arrays in torch layout: conv_w[i] = [Cout,Cin,3,3], conv_b[i] = [Cout]
 * (state_dict keys encoder.{0,1,3,4,6,7,9,10,12,13}.0.{weight,bias}); heads [16,2048] / [16]
 * (mu.0.*, logvar.0.*), flatten order c*16 + h*4 + w (latice/model.py:57).
 * ------------------------------------------------------------------------------------------- */
typedef struct ebsd_encoder ebsd_encoder;

typedef struct ebsd_weights {
    const float *conv_w[EBSD_N_CONV];
    const float *conv_b[EBSD_N_CONV]; /* accepted for layout parity; mathematically cancelled by InstanceNorm */
    const float *mu_w;
    const float *mu_b;
    const float *logvar_w;
    const float *logvar_b;
} ebsd_weights;

/* Packs the weights into the kernels' layouts (device-side; `w` arrays may be freed afterwards). */
int ebsd_encoder_create(ebsd_encoder **out, const ebsd_weights *w, int device, void *stream);
void ebsd_encoder_destroy(ebsd_encoder *enc);
/* Bytes of scratch ebsd_encoder_forward needs for a batch of B patterns. */
size_t ebsd_encoder_workspace_bytes(const ebsd_encoder *enc, int64_t B);
/* patterns: [B,128,128] (dtype EBSD_PATTERN_*); mu, logvar: [B,16] fp32 (logvar may be NULL). */
int ebsd_encoder_forward(ebsd_encoder *enc, const void *patterns, int dtype, int64_t B, float *mu, float *logvar,
                         void *workspace, size_t workspace_bytes, void *stream);

/* One encoder block alone (block `layer` = 1..9 of latice/model.py:109-125; block 1 includes conv0): what the
 * per-block parity tests and bench.py's per-block roofline table call.
 * layer 1: src = patterns [nimg,128,128] (dtype EBSD_PATTERN_*); src_sums [nimg,32,2] is scratch (conv0 statistics).
 * layers 2..9: src = raw fp32 NHWC [nimg,W,W,Cin] and src_sums [nimg,Cin,2] its plane sums over src_plane pixels;
 * the block applies InstanceNorm + LeakyReLU to src, convolves, and returns raw [nimg,Wo,Wo,Cout] (2x2 max-pooled
 * for layers 1,3,5,7,9) plus the plane sums [nimg,Cout,2] of the un-pooled output. */
int ebsd_encoder_block(ebsd_encoder *enc, int layer, int dtype, const void *src, double *src_sums, int src_plane,
                       int nimg, float *raw, double *sums, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Latent dictionary: exact cosine top-k
 *   replaces FaissLatentVectorDatabase._l2_normalize / add_vectors / query_similar
 *   (latice/index/faiss_db.py:109-113, 161-193, 216-256) and ChromaLatentVectorDatabase.query_similar
 *   (latice/index/chroma_db.py:231-259; cosine space, chroma_db.py:127-130).
 * Canonical arithmetic and tie order: see oracle/topk_ref.c.
 * ------------------------------------------------------------------------------------------- */
/* In place: x[i,:] /= ||x[i,:]|| (zero norm -> unchanged). d must be 16. */
int ebsd_normalize_rows(float *x, int64_t n, int d, void *stream);

size_t ebsd_topk_workspace_bytes(int64_t N, int64_t Q, int k);
/* dict: [N,16] normalised rows of this shard (16-byte aligned); global index of row r is index_base + r.
 * queries: [Q,16] normalised. Outputs [Q,k]: out_dot (q.d, descending), out_idx (global, -1 where N < k),
 * out_dist = 1 - dot (nullable). Order: dot desc, then global index asc. 1 <= k <= EBSD_MAX_TOPK. */
int ebsd_topk(const float *dict, int64_t N, int64_t index_base, const float *queries, int64_t Q, int k,
              float *out_dot, int64_t *out_idx, float *out_dist, void *workspace, size_t workspace_bytes,
              void *stream);
/* k-way merge of R per-shard candidate lists dots/idx [R,Q,k] (e.g. the all-gathered outputs of ebsd_topk). */
int ebsd_topk_merge(const float *dots, const int64_t *idx, int R, int64_t Q, int k, float *out_dot,
                    int64_t *out_idx, float *out_dist, void *stream);

/* The same exchange with one 64-bit word per candidate, (float bits of dot << 32) | (global row + 1) with 0 = empty
 * slot, so that the per-shard lists cross NVLink in ONE collective (needs fewer than 2^32 - 1 dictionary rows). */
int ebsd_topk_pack(const float *dots, const int64_t *idx, int64_t n, uint64_t *packed, void *stream);
int ebsd_topk_merge_packed(const uint64_t *packed, int R, int64_t Q, int k, float *out_dot, int64_t *out_idx,
                           float *out_dist, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Orientation consensus
 *   replaces find_best_orientation + _find_symmetry_equivalent_orientation
 *   (latice/index/chroma_db.py:261-375; FAISS twin latice/index/faiss_db.py:258-393),
 *   symmetry table latice/utils/constants.py:13-39.
 * ------------------------------------------------------------------------------------------- */
/* euler_deg [n,3] float64 (phi1, Phi, phi2; scipy "zxz" extrinsic, degrees) -> quat [n,4] float64 (x,y,z,w). */
int ebsd_euler_to_quat(const double *euler_deg, int64_t n, double *quat, void *stream);
/* quat_table [N,4] (32-byte aligned) / euler_table [N,3] (nullable): orientation of dictionary rows index_base ..
 * index_base + N - 1; cand_idx [Q,k]: global rows from ebsd_topk (-1 = empty).
 * Outputs per query: mean_quat [Q,4], mean_euler_deg [Q,3] (NaN when !success), success [Q],
 * similar_mask [Q] (bit i = candidate i within threshold in the last iteration run), ref_iter [Q], and (nullable)
 * cand_euler_deg [Q,k,3] = the stored Euler triplets of the candidates (OrientationResult.candidate_orientations,
 * chroma_db.py:283-289; NaN in empty slots). */
int ebsd_consensus(const double *quat_table, const double *euler_table, int64_t N, int64_t index_base,
                   const int64_t *cand_idx, int64_t Q, int k, double threshold, int angle_unit,
                   int min_required_matches, int max_iterations, int faiss_semantics, double *mean_quat,
                   double *mean_euler_deg, uint8_t *success, uint64_t *similar_mask, int32_t *ref_iter,
                   double *cand_euler_deg, void *stream);

/* ---------------------------------------------------------------------------------------------
 * IPF colour key of orientations (orientation maps)
 *   replaces get_color_key (latice/utils/utils.py:206-240) + ColorKeyGenerator.generate_ipf_color
 *   (latice/utils/colorkey.py:64-130).
 * euler_deg [n,3] float64 (scipy "zxz" extrinsic, degrees); axis 0 / 1 / 2 = mode "ipf_x" / "ipf_y" / "ipf_z" (the row
 * of the rotation matrix used as pole); rgb [n,3] uint8.
 * ------------------------------------------------------------------------------------------- */
int ebsd_ipf_color(const double *euler_deg, int64_t n, int axis, uint8_t *rgb, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EBSD_B200_H */
