/* sucre_b200 — C ABI of the B200-native SUCRe hot path (libsucre_b200.so).
 *
 * The reference (clementinboittiaux/sucre) has no FFI of its own: its hot path is a chain of PyTorch ATen
 * calls issued from Python.  Each entry point below replaces a span of reference Python, cited as file:line
 * into /root/reference/sucre/.  The intended binding is ctypes from the Python replacement of
 * sucre.restore_image (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; sucre_last_error() then describes it
 *     (thread-local, valid until the next call on the same thread);
 *   - all pointers are DEVICE pointers unless the parameter name ends in `_host`;
 *   - the caller owns all memory; nothing here allocates, frees, or synchronises the device;
 *     work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - images are row-major; a target pixel has flat index p = v*width + u; a "tile" is
 *     SUCRE_TILE_PIXELS (32) consecutive flat pixels, tile k covering p in [32k, 32k+32) — one warp, one lane
 *     per pixel.
 *
 * Observation store produced by the gather and consumed by the fit ("tile-major ELL rows"):
 *   `cells` is an array of ROWS of 32 records, one record per lane of a tile.  Within tile k, row j holds for lane
 *   i the j-th observation of the target pixel in SLOT 32k+i (its kept source views in pairing-list order) or, when
 *   the pixel has fewer than j+1 observations, an all-zero sentinel record (a real observation always has z > 0).
 *   The pixel in a slot is pixel 32k+i itself, or, after sucre_gather_permute, the one the store's `pix` map names
 *   (struct sucre_store below): the pixels of 32 consecutive tiles dealt to the slots by observation count, which
 *   leaves few sentinels.  Tile k has
 *   rows(k) = max over its 32 pixels of the observation count and starts at row row_off[k]; tiles follow each other,
 *   so any run of rows is one contiguous byte range (the fit streams it with 1-D TMA bulk copies), a warp reads a
 *   row as one fully coalesced, bank-conflict-free access, and there are no headers, lane offsets or segment
 *   boundaries to decode.  Record formats (record_format):
 *     SUCRE_REC_Z_U8   8 bytes  {float z; uint8 r, g, b, 0}      z = ||cP||, the range of the observation in the
 *                               source camera frame (loader.py:113 + sucre.py:53); r,g,b the source pixel's u8 colour
 *                               (the reference's I = u8 / 255, loader.py:157, 87, is formed in the fit's arithmetic)
 *     SUCRE_REC_Z_F32  16 bytes {z, I_r, I_g, I_b} floats        scenes whose colour was resampled in float
 *                               (--image-scale, loader.py:158-162)
 *     SUCRE_REC_P_U8   16 bytes {cP_x, cP_y, cP_z; uint8 r, g, b, 0}   the light model (--light-model) needs the
 *                               camera-frame point itself (sucre.py:57)
 *     SUCRE_REC_P_F32  32 bytes {cP_x, cP_y, cP_z, ||cP||} {I_r, I_g, I_b, 0}
 *   Bytes per row = 32 * record bytes.  Total rows = row_off[n_tiles]; fill = N / (32 * rows).
 *   For export and parity checks the gather also lists, per tile, its non-empty kept (tile, view) blocks in
 *   pairing-list order: blk_mask[b] = lane mask, blk_view[b] = view index, b in blk_off[k]..blk_off[k+1]; the j-th
 *   record of lane i is the j-th block of the tile whose mask has bit i set.
 */
#ifndef SUCRE_B200_H
#define SUCRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUCRE_ABI_VERSION 10
#define SUCRE_TILE_PIXELS 32
#define SUCRE_REC_Z_U8 0
#define SUCRE_REC_Z_F32 1
#define SUCRE_REC_P_U8 2
#define SUCRE_REC_P_F32 3

/* One view (source or target).  All matrices row-major fp32, computed on the host with the reference's own
 * expressions so that they are bit-identical to what the reference multiplies by:
 *   K     Camera.K                        sfm.py:204-208
 *   Kinv  K.inverse()                     sfm.py:92
 *   R, t  cam->world Pose                 sfm.py:219-222 (inverse of COLMAP's cam_from_world)
 *   Ri,ti Pose.inverse(): R.T, -R.T @ t   sfm.py:47
 * depth: u16 millimetres, height*width (loader.py:167 divides by 1000 -> metres; done in-kernel, IEEE).
 * rgb:   rgb_format SUCRE_RGB_U8:  u8 RGB interleaved, height*width*3 (loader.py:157 divides by 255; done in-kernel,
 *                                      IEEE) — images used at their stored size;
 *        rgb_format SUCRE_RGB_F32: float RGB interleaved in [0,1], height*width*3 — images the host resampled with
 *                                      --image-scale (loader.py:158-162 resamples in float, which leaves the u8 grid).
 * sizeof == 208, 16-byte aligned. */
#define SUCRE_RGB_U8 0
#define SUCRE_RGB_F32 1
/* K (resp. Kinv) has the PINHOLE sparsity [[a,0,b],[0,c,d],[0,0,1]] exactly (entries 1,3,6,7 are +-0 and entry 8 is 1):
 * the kernels then skip the multiplications by 0 and 1, which cannot change any rounded result (x*0 + y == y). */
#define SUCRE_VIEW_K_SPARSE 1
#define SUCRE_VIEW_KINV_SPARSE 2
typedef struct sucre_view {
    float K[9], Kinv[9], R[9], t[3], Ri[9], ti[3];
    int32_t width, height;
    const uint16_t* depth;
    const void* rgb;
    int32_t rgb_format;
    int32_t flags;           /* SUCRE_VIEW_* bits, set by the host from the VALUES of K / Kinv (never assumed) */
    int32_t reserved[2];
} sucre_view;

/* The observation store of one target (or of a band of its tiles), as the fit reads it.  A host struct of
 * device pointers.  sizeof == 48.
 *
 * Slots and pixels.  Lane i of tile k is SLOT 32k + i.  Without `pix`, slot q is the store's q-th pixel (valid iff
 * q < pixels).  With `pix` (sucre_gather_permute), the pixels of every group of 32 tiles were dealt to the group's slots
 * in order of decreasing observation count — neighbouring lanes then have columns of nearly equal height and the ELL
 * rows hold few sentinels (fill 0.90 -> 0.99 on the benchmark scene) — and pix[q] is the pixel of slot q (-1: none).
 * The J / J_moments arrays of the Adam-loop entry points (sucre_fit, sucre_fit_sharded, sucre_fit_sums) and J_ref of
 * sucre_fit_write_J are always in SLOT order (32 * n_tiles entries when pix is set, `pixels` otherwise): they are work
 * buffers of the loop.  sucre_fit_write_J's output, the light-model entry points and sucre_band_scatter_J address J by
 * PIXEL (pix[q]; `pixels` = entries of those arrays). */
typedef struct sucre_store {
    const void* cells;       /* rows of 32 records, 16-byte aligned; may be NULL when n_rows == 0 */
    const int64_t* row_off;  /* [n_tiles+1] rows before tile k */
    int32_t n_tiles;
    int32_t record_format;   /* SUCRE_REC_* */
    int64_t pixels;          /* pix == NULL: target pixels covered, min(n_tiles*32, width*height - first_tile*32);
                                pix != NULL: entries of the pixel-ordered arrays pix indexes (e.g. width*height) */
    int64_t n_rows;          /* host copy of row_off[n_tiles] */
    const int32_t* pix;      /* [n_tiles*32] pixel of every slot, -1 = none; NULL = the identity */
} sucre_store;

/* The tiles of a target that one call (one rank) covers.  Local tile k is the target's tile
 *   first_tile + (k / chunk_tiles) * stride_tiles + k % chunk_tiles,      k in [0, n_tiles).
 * Whole image or contiguous band: chunk_tiles = n_tiles.  Cyclic sharding over N ranks in chunks of C tiles (every
 * rank then sees every region of the image, so the observation counts balance): rank r has first_tile = r*C,
 * chunk_tiles = C, stride_tiles = N*C.  masks / offsets / cells / J of a call are in LOCAL tile order. */
typedef struct sucre_band {
    int32_t first_tile, n_tiles, chunk_tiles, stride_tiles;
} sucre_band;

int sucre_abi_version(void);
int sucre_record_bytes(int record_format); /* 8, 16, 16, 32; 0 for an unknown format */
const char* sucre_last_error(void);

/* ---- scene upload ---------------------------------------------------------------------------------------------
 * Replaces the per-(target, view) decode + `.to(device)` of whole source images (sfm.py:130-133 via
 * loader.py:156-170): copies, for n views, only the rectangle of each view that the target can see
 * (engine.DeviceScene.footprints: a conservative bound of where target pixels can land) from a host stack of
 * equally sized planes into the matching device stack, with cudaMemcpy2DAsync on `stream`.
 *   dst            device stack: view i starts at dst + i*view_bytes
 *   src_host       host stack (pinned for asynchronous copies): view src_index_host[i] starts at
 *                  src_host + src_index_host[i]*view_bytes
 *   width, height, pixel_bytes   plane geometry: view_bytes = height * width * pixel_bytes (2: u16 depth, 3: u8 RGB)
 *   rects_host     int32[n][4] = {x0, y0, x1, y1} in pixels, half-open; x1 <= x0 or y1 <= y0: nothing copied
 *   copied_bytes_host  (optional) receives the bytes enqueued
 * Pixels outside the rectangles are left as they are (the caller zero-fills the device stack: depth 0 = invalid,
 * sfm.py:96). */
int sucre_scene_upload(void* dst, const void* src_host, int n, const int32_t* src_index_host, int width, int height,
                       int pixel_bytes, const int32_t* rects_host, int64_t* copied_bytes_host, void* stream);

/* ---- stage 1: multi-view correspondence gather ---------------------------------------------------------
 * Replaces Image.match_images / match_two_way / match_one_way / Matches.map / Matches.__and__
 * (sfm.py:115-138, 154-159, 171-175), unproject_depth(_map) / project_to_view / Pose.transform
 * (sfm.py:49-55, 90-107) and load_depth_map's scaling (loader.py:166-170).
 *
 * A call may cover the whole target or a band of it (struct sucre_band, multi-GPU pixel sharding); masks / offsets /
 * cells / J of such a call are local to the band.
 *
 * sucre_gather_match: for every target pixel of the band and every listed view, the reference's two-way integer
 * round-trip test.  masks[k*n_views + s] receives the lane mask of local tile k against view s.
 * Bit-exact with the reference (SURVEY.md §8a').  A conservative per-(warp, view) frustum test skips views in which
 * none of the warp's 64 target pixels can land (their mask words are written as 0, which is what the full evaluation
 * gives).  stats (optional, may be NULL; int64[2], ACCUMULATED, zero it first): [0] += (tile, view) pairs skipped by
 * that test, [1] += forward projections that landed inside the source image (SURVEY.md §8d's n_inbounds). */
int sucre_gather_match(const sucre_view* target_host, const sucre_view* views, int n_views, const sucre_band* band_host,
                       uint32_t* masks, int64_t* stats, void* stream);

/* sucre_gather_permute (optional, between sucre_gather_match and sucre_gather_plan): deals the pixels of every group
 * of SUCRE_GROUP_TILES consecutive local tiles to the group's slots in order of decreasing match count (over all listed
 * views; ties in pixel order, so the result is deterministic).
 *   pix[n_tiles*32]          (int32) flat index IN THE TARGET IMAGE of the pixel in every slot, -1 for none
 *   pmasks[n_tiles*n_views]  the masks of the permuted tiles: bit i of pmasks[k*n_views + s] = pixel pix[32k+i] matched
 *                            in view s.  Hand pmasks (not masks) to sucre_gather_plan and sucre_gather_sample, and pix
 *                            to sucre_gather_sample and to the store. */
#define SUCRE_GROUP_TILES 32
int sucre_gather_permute(const uint32_t* masks, int n_views, const sucre_band* band_host, int64_t target_pixels,
                         int32_t* pix, uint32_t* pmasks, void* stream);

/* sucre_gather_count: view_count[n_views] (int64) = matches per view over the band (kept or not).
 * Multi-GPU callers all-reduce view_count before sucre_gather_plan: min_cover is a whole-image criterion. */
int sucre_gather_count(const uint32_t* masks, int n_tiles, int n_views, int64_t* view_count, void* stream);

/* sucre_gather_plan: the min_cover decision (sfm.py:136: a view is kept iff count / (width*height) > min_cover,
 * evaluated in double like the reference's Python floats; view_count and target_pixels are WHOLE-IMAGE figures)
 * and the layout of the band's observation store.
 *   view_kept[n_views]   (uint8)  1 if the view passes min_cover
 *   rec_off, blk_off, row_off [n_tiles+1] (int64) exclusive prefix sums over the band's tiles, kept views only:
 *                        records (observations), blocks = non-empty (tile, view) pairs, rows = the largest observation
 *                        count among the tile's 32 pixels
 *   totals[3] (int64)    {N = observations in the band, blocks, rows}; copy to the host to size the store:
 *                        cells = rows * 32 * sucre_record_bytes(format) bytes, blk_mask / blk_view = blocks entries */
int sucre_gather_plan(const uint32_t* masks, int n_tiles, int n_views, const int64_t* view_count,
                      int64_t target_pixels, double min_cover, uint8_t* view_kept, int64_t* rec_off, int64_t* blk_off,
                      int64_t* row_off, int64_t* totals, void* stream);

/* sucre_gather_sample: fills the observation store.  Replaces MatchesFile.save_matches / prepare_matches /
 * load_matches (loader.py:68-87, 103-118) and load_rgb's scaling (loader.py:156-163); the HDF5 spill file is
 * replaced by this device-resident store.  record_format: SUCRE_REC_*; the U8 formats require every kept view to be
 * SUCRE_RGB_U8.  cell_src (optional, may be NULL; one uint32 per record slot, index row*32 + lane) receives
 * u2 | v2 << 16, the integer source pixel (what the reference stores as int16 u2, v2), 0xffffffff at sentinels.
 * pix (optional, may be NULL): the slot -> pixel map of sucre_gather_permute; masks are then its pmasks. */
int sucre_gather_sample(const sucre_view* target_host, const sucre_view* views, int n_views, const sucre_band* band_host,
                        const int32_t* pix, const uint32_t* masks, const uint8_t* view_kept, const int64_t* row_off,
                        const int64_t* blk_off, int record_format, void* cells, uint32_t* blk_mask, int32_t* blk_view,
                        uint32_t* cell_src, void* stream);

/* sucre_band_scatter_J: copies a band's J (J_band[n_tiles*32*3], slot order) to its place in n_dst whole-image
 * J buffers (target_pixels*3 floats each; dst_ptrs_host[i] = device address — this GPU's or a peer's NVLink-mapped
 * buffer).  pix (optional): the band's slot -> image pixel map; NULL = the band's own tile arithmetic.
 * The assembly step of a target sharded over several GPUs ("the final gather of J") as direct peer writes. */
int sucre_band_scatter_J(const float* J_band, const sucre_band* band_host, const int32_t* pix, int64_t target_pixels,
                         const uint64_t* dst_ptrs_host, int n_dst, void* stream);

/* ---- stage 2: per-pixel fit of the image formation model ------------------------------------------------
 * Replaces SUCRe.compute_l_z / update_J / forward (sucre.py:52-82) and adam() (sucre.py:124-157) for
 * light_model=False.  params = {B[3], beta[3], gamma[3]} fp32 (sucre.py:41-43).
 *
 * mode SUCRE_FIT_CLOSED_FORM  --use-closed-form: per pixel J = sum((I - B(1-e^{-gamma z})) e^{-beta z}) /
 *                             sum(e^{-2 beta z}) from the CURRENT params (sucre.py:66-77), residuals against it.
 *                             J[pixels*3] is a work buffer (zero it before the first iteration): it carries the
 *                             J of the previous iteration, the reference point of the single-sweep statistics.
 * mode SUCRE_FIT_PARAM_J      default CLI mode: J[pixels*3] is an Adam parameter (sucre.py:47-50), initialised by
 *                             the caller to the target image with NaN where target depth <= 0; J_moments
 *                             [pixels*6] = per pixel {exp_avg[3], exp_avg_sq[3]}, zero-initialised.
 * Every iteration reads each row exactly once.  Stores must be SUCRE_REC_Z_U8 or SUCRE_REC_Z_F32. */
#define SUCRE_FIT_CLOSED_FORM 0
#define SUCRE_FIT_PARAM_J 1

/* Size of the scratch buffer of the fit calls (16-byte aligned, one per observation store). */
size_t sucre_fit_workspace_bytes(void);

/* Once per observation store, before any other fit call on `workspace`: partitions the rows over the resident
 * warps (equal rows + per-tile overhead per warp; a tile may be split between two neighbouring warps of a CTA, which
 * combine their per-pixel statistics through shared memory; static => reproducible summation order). */
int sucre_fit_prepare(const sucre_store* store_host, void* workspace, void* stream);

/* One evaluation of the objective at `params`, reduced to sums[10] (double):
 *   sums[0..2] = sum r(1-e^{-gamma z}), sums[3..5] = sum r J z e^{-beta z}, sums[6..8] = sum r B z e^{-gamma z},
 *   sums[9] = sum r^2 (the `cost` the reference logs, sucre.py:144-146,150),
 * r = I - (J e^{-beta z} + B(1-e^{-gamma z})) (sucre.py:81).  In SUCRE_FIT_PARAM_J mode the per-pixel Adam step t of
 * J (gradient -(2/(3 n_obs)) sum r e^{-beta z}) is applied in the same pass; n_obs, t, lr are ignored otherwise.
 * Multi-GPU callers all-reduce sums between this call and sucre_adam_step; n_obs is then the global count. */
int sucre_fit_sums(int mode, const sucre_store* store_host, const float* params, float* J, float* J_moments,
                   int64_t n_obs, int t, double lr, double* sums, void* workspace, void* stream);

/* Adam step t (1-based) on the 9 parameters from the reduced sums: gradients of sum r^2 / (3 n_obs)
 * (sucre.py:144-145), update rule of torch.optim.Adam defaults (sucre.py:136,148: betas .9/.999, eps 1e-8,
 * fp32 state).  adam_state = {m[9], v[9]}.  history_row (optional) receives {params after the step [9], cost}. */
int sucre_adam_step(float* params, float* adam_state, const double* sums, int64_t n_obs, int t, double lr,
                    float* history_row, void* stream);

/* The whole single-GPU loop of adam() (sucre.py:138-148) as ONE launch of a resident grid (one per 1024 iterations):
 * every iteration = one sweep of the store, after which every CTA publishes its row of partial sums, gathers all rows
 * and takes the Adam step of the 9 scalars on its own copy of the state (steps first_step .. first_step+num_iter-1).  history (optional) = num_iter x 10
 * floats.  For the final update_J of closed-form mode (sucre.py:156) call sucre_fit_write_J. */
int sucre_fit(int mode, const sucre_store* store_host, int64_t n_obs, float* params, float* adam_state, float* J,
              float* J_moments, int first_step, int num_iter, double lr, float* history, void* workspace,
              void* stream);

/* The same loop for ONE target whose tiles are sharded over `world` GPUs (store_host = this rank's band, n_obs_global
 * = observations of all bands): the all-reduce of the 10 sums is fused into the kernel — CTA 0 of every rank
 * stores its rank's sums, as 8-byte words tagged with the iteration's epoch, into every peer's exchange buffer over NVLink
 * (peers_host[p] = device address of rank p's buffer, SUCRE_PEER_BUFFER_BYTES each, zeroed once, mapped into this
 * process, e.g. torch symmetric memory), every CTA polls its own GPU's buffer for the words of all ranks and adds the rows in rank
 * order, so every rank applies the identical Adam step with no host or NCCL round trip.  first_epoch: a tag >= 1 for the first iteration, identical on all ranks, and increasing by num_iter
 * from one call on the same buffers to the next.  All ranks must make the same sequence of calls.  A rank that waits
 * longer than SUCRE_PEER_TIMEOUT_NS for a peer stops waiting, sets bit 0 of the status word
 * (sucre_fit_status) and carries on with what it has, so a dead peer cannot hang the node. */
#define SUCRE_MAX_PEERS 16
#define SUCRE_PEER_BUFFER_BYTES 5120
#define SUCRE_PEER_TIMEOUT_NS 2000000000ull
int sucre_fit_sharded(int mode, const sucre_store* store_host, int64_t n_obs_global, float* params, float* adam_state,
                      float* J, float* J_moments, int first_step, int num_iter, double lr, float* history,
                      void* workspace, const uint64_t* peers_host, int rank, int world, uint32_t first_epoch,
                      void* stream);

/* Copies the status word of `workspace` (0 = fine; bit 0: a peer exchange timed out) to *status (device
 * pointer, uint32) on `stream`, and clears it. */
int sucre_fit_status(void* workspace, uint32_t* status, void* stream);

/* Closed-form J for the current params written to J[pixels*3] in PIXEL order (through store->pix when set); NaN where
 * a pixel has no observation (0/0 like sucre.py:77).  J_ref (optional): a previous J (the loop's work buffer, SLOT
 * order) used as the reference point of the statistics. */
int sucre_fit_write_J(const sucre_store* store_host, const float* params, const float* J_ref, float* J,
                      void* workspace, void* stream);

/* ---- stage 2, light model (--light-model, sucre.py:44-46, 54-61) ---------------------------------------------
 * l = exp(-lp^T Sigma^-1 lp / 2) with lp the perspective division of lP = R cP + t, z = ||cP|| + ||lP||,
 * I_hat = l (J e^{-beta z} + B (1 - e^{-gamma z})).  Stores must be SUCRE_REC_P_U8 or SUCRE_REC_P_F32.
 * params24 = B[3], beta[3], gamma[3], R[9] row-major, t[3], Sinv[3] = (Sigma^-1)_00, _01, _11 — R, t =
 * se3.exp(cam2light) (se3.py:22-27) and Sigma^-1 = (sigma^T sigma)^-1 are evaluated on the host with the reference's
 * own torch expressions; the chain rule back to cam2light / sigma is applied on the host from sums[10..24].
 *
 * sucre_light_J: closed-form J (sucre.py:66-77 with the light terms) for the given parameters; NaN where unobserved. */
int sucre_light_J(const sucre_store* store_host, const float* params24, float* J, void* stream);

/* sucre_light_sums: one residual pass with J given per pixel, reduced to sums[25] (double):
 *   [0..2] sum r l (1-e^{-gamma z})   [3..5] sum r l J z e^{-beta z}   [6..8] sum r l B z e^{-gamma z}   [9] sum r^2
 *   [10..12] sum g_l l {x^2, x y, y^2}           (lp = (x, y);  dL/dSigma^-1 = -(2/3N) * (-1/2) * [[.,.],[.,.]])
 *   [13..21] sum g_lP (x) cP (3x3 row-major)      (dL/dR = -(2/3N) * this)
 *   [22..24] sum g_lP                             (dL/dt = -(2/3N) * this)
 * with g_l = sum_c r_c (J_c a_c + B_c h_c), g_z = sum_c r_c l (-beta_c J_c a_c + gamma_c B_c g_c),
 * g_lP = g_l dl/dlP + g_z lP/||lP||.  mode SUCRE_FIT_PARAM_J additionally applies Adam step t to J in the same pass
 * (gradient -(2/(3 n_obs)) sum r l e^{-beta z}); in SUCRE_FIT_CLOSED_FORM mode J is read only. */
int sucre_light_sums(int mode, const sucre_store* store_host, const float* params24, float* J, float* J_moments,
                     int64_t n_obs, int t, double lr, double* sums, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUCRE_B200_H */
