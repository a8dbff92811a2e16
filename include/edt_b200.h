/* include/edt_b200.h -- C ABI of the B200-native multi-label Euclidean distance transform.
 *
 * This is the drop-in boundary for the reference's hot path
 *     pyedt::_edt3dsq<T>(labels, sx,sy,sz, wx,wy,wz, black_border, parallel, workspace)
 *                                                    (reference src/edt.hpp:411-484)
 * and its 2-D / 1-D siblings (src/edt.hpp:632-678, 70-119), which the reference's Cython
 * layer binds in src/edt.pyx:63-87 (`cdef extern from "edt.hpp" namespace "pyedt"`).
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Conventions shared by every entry point
 *   - volumes are x-fastest ("Fortran order" in the reference's words, README.md:100);
 *     C-ordered numpy arrays are handled by the caller swapping axes (src/edt.pyx:651-664);
 *   - labels are compared for equality only, as raw 1/2/4/8-byte unsigned integers
 *     (the reference instantiates uint8/16/32/64, float, double and bool: src/edt.pyx:670-732;
 *     float labels must be canonicalised by the caller so that -0.0 == +0.0);
 *   - the output is float32, same shape; with EDTB200_OUT_ON_DEVICE the transform runs in
 *     place in `out` (the reference's `workspace` argument has the same role);
 *   - every function returns 0 on success or a negative EDTB200_E* code;
 *     edtb200_last_error() gives the message (thread-local).  Nothing throws.
 *   - there is NO CPU fallback: without a usable CUDA device every transform fails with
 *     EDTB200_ECUDA.
 */
#ifndef EDT_B200_H
#define EDT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 100: round 1 (transform, per-axis passes, slab face kernels).  200: additions only --
 * edtb200_transform_multi, edtb200_slab_step / _stage_bytes / _pack, edtb200_label_stats / _extract,
 * edtb200_host_alloc / _free; every 100-level entry point keeps its signature and meaning. */
#define EDTB200_VERSION 200

/* flags */
#define EDTB200_SQRT             1  /* emit sqrt(edtsq): reference _edt3d (src/edt.hpp:591-604) /
                                       np.sqrt in edt() (src/edt.pyx:241-242), fused in the last store */
#define EDTB200_SIGNED           2  /* signed distance: background (label 0) is transformed as an
                                       ordinary label and its result negated -- equal to the
                                       reference's sdf()/sdfsq() = f(data) - f(data==0)
                                       (src/edt.pyx:121-202) */
#define EDTB200_LABELS_ON_DEVICE 4  /* `labels` is a device pointer on `device` */
#define EDTB200_OUT_ON_DEVICE    8  /* `out` is a device pointer on `device` */
#define EDTB200_LABELS_FLOAT    16  /* edtb200_transform_voxel_graph only: the 4- / 8-byte labels are
                                       IEEE floats and foreground means value > 0, as the reference's
                                       float instantiations test it (src/edt_voxel_graph.hpp:76, 151) */

/* error codes */
#define EDTB200_EINVAL  (-1)   /* bad argument */
#define EDTB200_ECUDA   (-2)   /* CUDA runtime / driver error, or no device */
#define EDTB200_ENOMEM  (-3)   /* device or host allocation failed */
#define EDTB200_ELIMIT  (-4)   /* size outside what the kernels support */

int edtb200_version(void);
const char *edtb200_last_error(void);

/* Number of visible CUDA devices (0 if none or on error). */
int edtb200_device_count(void);

/* Whole transform: edtsq / edt / sdfsq / sdf of a 1-D, 2-D or 3-D volume.
 *
 * Replaces  pyedt::_edt3dsq<T>        src/edt.hpp:411-484   (ndim == 3)
 *           pyedt::_edt2dsq<T>        src/edt.hpp:632-678   (ndim == 2; sz ignored)
 *           squared_edt_1d_multi_seg  src/edt.hpp:70-119    (ndim == 1; sy, sz ignored)
 *           pyedt::_edt3d / _edt2d    src/edt.hpp:591-604, 764-778  (EDTB200_SQRT)
 *           edt.sdf / edt.sdfsq       src/edt.pyx:121-202   (EDTB200_SIGNED)
 * For ndim < 3 the missing axis passes are skipped (not run on size-1 axes), exactly as the
 * reference's lower-dimensional drivers do.
 *
 *   labels       sx*sy*sz labels of `label_bytes` (1, 2, 4 or 8) bytes each; never written
 *   wx, wy, wz   anisotropy (voxel size) per axis
 *   black_border non-zero: the volume faces count as background
 *   out          sx*sy*sz float32
 *   device       CUDA device ordinal
 *   stream       cudaStream_t (as void*) or NULL.  Device-resident calls (both *_ON_DEVICE
 *                flags) are asynchronous on `stream`; calls touching host memory return
 *                after the result has landed in `out`.
 */
int edtb200_transform(const void *labels, int label_bytes, int ndim,
                      int64_t sx, int64_t sy, int64_t sz,
                      float wx, float wy, float wz,
                      int black_border, int flags,
                      float *out, int device, void *stream);

/* The same transform of ONE host volume spread over several GPUs of this process
 * (`devices[0..ndevices)`, distinct ordinals): SURVEY.md section 8b's `devices[] / ndevices`.
 * Every device uploads one Z slab over its own PCIe link and runs the X and Y passes on it; the
 * distances and labels are then re-partitioned into Y slabs (whole z lines per device) by
 * peer-to-peer 3-D copies over NVLink, the Z pass runs, the result goes back the same way and
 * down to `out`.  Exact for any input (no halo, no verdict); `labels` and `out` are HOST pointers
 * (the *_ON_DEVICE flags are rejected).  Volumes that are not 3-D, or too thin to split, run on
 * devices[0].  Device memory per GPU: 2 * (label_bytes + 4) * voxels / ndevices.  Calls on
 * disjoint device sets may run concurrently from different threads. */
int edtb200_transform_multi(const void *labels, int label_bytes, int ndim,
                            int64_t sx, int64_t sy, int64_t sz,
                            float wx, float wy, float wz,
                            int black_border, int flags,
                            float *out, const int *devices, int ndevices);

/* Many volumes of one geometry, HOST buffers, pipelined: while volume k is being transformed,
 * volume k+1 is on its way to the device and volume k-1 on its way back (two device slots,
 * separate copy streams, PCIe used in both directions at once).  This is what a caller of the
 * reference does in a loop over the chunks of a dataset (README.md:191, the skeletonisation use
 * case); per volume the result is exactly what edtb200_transform returns.
 *   labels[k], outs[k]  host pointers (pinned memory is copied directly, pageable memory is staged
 *                       by two groups of host threads); outs[k] must not alias any labels[j]
 *   flags               EDTB200_SQRT / EDTB200_SIGNED; the *_ON_DEVICE flags are rejected
 * Returns when every result has landed. */
int edtb200_transform_batch(const void *const *labels, float *const *outs, int count,
                            int label_bytes, int ndim, int64_t sx, int64_t sy, int64_t sz,
                            float wx, float wy, float wz, int black_border, int flags, int device);

/* Transform under a voxel connectivity graph (2-D or 3-D only).
 *
 * Replaces  pyedt::_edt2dsq_voxel_graph<T, uint8_t>   src/edt_voxel_graph.hpp:54-123
 *           pyedt::_edt3dsq_voxel_graph<T, uint8_t>   src/edt_voxel_graph.hpp:125-214
 *           pyedt::_edt3d_voxel_graph  (EDTB200_SQRT) src/edt_voxel_graph.hpp:216-236
 * as bound by __edt2dsq_voxel_graph / __edt3dsq_voxel_graph, src/edt.pyx:514-620, 736-844.
 *
 *   graph   sx*sy*sz bytes, laid out like the labels; bit 0 / 2 / 4 set = the step to the +x / +y /
 *           +z neighbour is allowed (the cc3d voxel-connectivity bit field; other bits are not read)
 * Labels only say foreground (non-zero; > 0 with EDTB200_LABELS_FLOAT) or background here, as in
 * the reference.  The foreground is drawn on a doubled grid in which a forbidden edge is a
 * background cell, transformed with half the anisotropy, and sampled back; all of it on the device
 * (scratch: 1 + 4 bytes per doubled cell).  EDTB200_LABELS_ON_DEVICE covers `labels` AND `graph`.
 * EDTB200_SIGNED is rejected: the reference's sdf with a graph is the difference of two such
 * transforms (src/edt.pyx:147-158), which the caller forms.
 */
int edtb200_transform_voxel_graph(const void *labels, int label_bytes, const unsigned char *graph,
                                  int ndim, int64_t sx, int64_t sy, int64_t sz,
                                  float wx, float wy, float wz,
                                  int black_border, int flags,
                                  float *out, int device, void *stream);

/* Single axis passes on DEVICE-resident data, for callers that split a volume into Z slabs
 * across GPUs: every rank runs the first- and second-axis passes on
 * its slab, exchanges what the third axis needs, then runs the third-axis pass.
 *
 * edtb200_pass_first : replaces the X loop of _edt3dsq          src/edt.hpp:430-440
 * edtb200_pass_later : replaces the Y or Z loop of _edt3dsq     src/edt.hpp:450-475
 *     axis 1 = Y (stride sx), axis 2 = Z (stride sx*sy).
 *     flags: on pass_first EDTB200_SIGNED means "background is an ordinary label" (needed for
 *     sdf, no sign is applied there); on pass_later EDTB200_SQRT / EDTB200_SIGNED are applied
 *     in that pass's store (sqrt, negate background) -- use them on the last pass only.
 *     border_lo / border_hi say independently whether the low / high end of the axis is a
 *     volume face with black_border (an interior slab face is neither).
 *     pass_later works IN PLACE on f_dev (input = the previous pass's output) and only writes
 *     the elements whose value it changes.
 */
int edtb200_pass_first(const void *labels_dev, int label_bytes,
                       int64_t sx, int64_t sy, int64_t sz, float wx,
                       int black_border, int flags,
                       float *f_dev, int device, void *stream);

int edtb200_pass_later(const void *labels_dev, int label_bytes, int axis,
                       int64_t sx, int64_t sy, int64_t sz, float w,
                       int border_lo, int border_hi, int flags,
                       float *f_dev, int device, void *stream);

/* Z-slab decomposition across GPUs: a rank runs edtb200_pass_later(axis=2)
 * on its slab with the interior faces open (border flags 0 there) and then folds the
 * neighbouring slabs in with these two calls, which replace what the reference gets for free
 * from having the whole z-line in one address space (src/edt.hpp:465-475).
 *
 * edtb200_slab_face_runs : for every (x,y) line, the length m (1..halo) of the run of equal
 *     labels that touches the low (high_face=0) or high (high_face=1) face of this slab;
 *     m = halo+1 means "longer than the halo, or spanning the whole slab".  *overflow_dev (device
 *     int, zeroed by the caller) is raised when that happens for a foreground run (any run with
 *     EDTB200_SIGNED): a hint that the fix-up's verdict below matters for this volume.
 * edtb200_slab_face_fixup : after the third-axis pass (same flags: EDTB200_SQRT / SIGNED),
 *     min-combines every row of the face-touching runs with the neighbour's sites:
 *     nb_label_dev = the neighbour's face plane of labels, nb_m_dev = its face_runs output,
 *     nb_f_dev = `halo` planes of the neighbour's distances taken after ITS second-axis pass,
 *     in the neighbour's z order (its last `halo` planes for our low face, its first for our
 *     high face).  Requires sz > halo on both sides.  Runs that go on behind the halo are exact
 *     as long as the distances at the face stay within the halo's reach (value <= (w*halo)^2);
 *     *inexact_dev (device int, zeroed by the caller, may be NULL) is raised when they do not,
 *     and the caller must then repeat with a deeper halo or an exact method.
 */
int edtb200_slab_face_runs(const void *labels_dev, int label_bytes,
                           int64_t sx, int64_t sy, int64_t sz, int high_face, int halo, int flags,
                           unsigned char *m_dev, int *overflow_dev, int device, void *stream);

int edtb200_slab_face_fixup(const void *labels_dev, int label_bytes,
                            int64_t sx, int64_t sy, int64_t sz, int high_face, int halo, float wz,
                            int flags, const void *nb_label_dev, const unsigned char *nb_m_dev,
                            const float *nb_f_dev, float *f_dev, int *inexact_dev,
                            int device, void *stream);

/* Exact fallback of the Z-slab decomposition (runs deeper than any halo): the ranks re-partition
 * the volume from Z slabs to Y slabs, run the third-axis pass on whole z-lines and re-partition
 * back.  edtb200_slab_pack moves a device-resident slab (zc, sy, row_bytes) -- distances or
 * labels, the row is opaque bytes -- into the exchange layout in ONE launch: the slab is cut
 * along y into `parts` pieces (piece i = rows y_start[i] .. y_start[i+1], y_start[0] = 0,
 * y_start[parts] = sy; HOST array of parts + 1 entries, parts <= 64), and piece i becomes one
 * contiguous block (zc, c_i, row_bytes), the blocks in piece order -- block i is the single
 * message for rank i.  unpack != 0 applies the inverse map (blocks -> slab) for the way back.
 * Replaces what the reference gets from having the whole volume in one address space
 * (src/edt.hpp:450-475).
 */
int edtb200_slab_pack(const void *src_dev, void *dst_dev, int64_t zc, int64_t sy, int64_t row_bytes,
                      int parts, const int64_t *y_start, int unpack, int device, void *stream);

/* One whole step of the Z-slab decomposition on this rank, device-resident, asynchronous on
 * `stream`: X pass, Y pass, publication of this slab's faces to the neighbours, Z pass (interior
 * faces open) and the fix-up that folds the neighbours' rows in -- five launches from ONE call, no
 * host synchronisation, no collective.  The ranks meet only through flag words in a staging buffer
 * that every rank allocates in CUDA symmetric memory (peer-mapped over NVLink):
 *   sym_self        this rank's buffer, edtb200_slab_stage_bytes(sx, sy, label_bytes, halo) bytes,
 *                   ZERO-FILLED once before the first step (all ranks, then a barrier)
 *   sym_lo, sym_hi  the lower / upper neighbour's buffer as mapped into this process (NULL where
 *                   has_lo / has_hi is 0)
 *   step            1, 2, 3, ... the same on every rank; staging sets alternate by its parity
 *   status_dev      device int, OR-ed with 1 when the halo was too shallow for this volume (a run
 *                   goes on behind the neighbour's `halo` rows and the distances at the face exceed
 *                   the halo's reach: repeat with a deeper halo or an exact method), and with 2
 *                   when a neighbour never published the step (time-out)
 * flags: EDTB200_SQRT / EDTB200_SIGNED as for edtb200_transform; black_border applies to the real
 * volume faces only (z low face on the rank with has_lo == 0, z high face where has_hi == 0).
 * This is what replaces the reference's single-address-space Z loop (src/edt.hpp:465-475) when
 * the volume is spread over several GPUs. */
int64_t edtb200_slab_stage_bytes(int64_t sx, int64_t sy, int label_bytes, int halo);

int edtb200_slab_step(const void *labels_dev, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
                      float wx, float wy, float wz, int black_border, int has_lo, int has_hi, int flags,
                      float *f_dev, int halo, void *sym_self, void *sym_lo, void *sym_hi,
                      unsigned long long step, int *status_dev, int device, void *stream);

/* Per-label views of a finished transform, on the device: the reference's `edt.each`
 * (src/edt.pyx:951-994: one masked distance image per label, built from run lists,
 * src/edt_voxel_graph.hpp:238-310) for labels and distances that are already resident on the GPU.
 *
 * edtb200_label_stats : ONE pass over (labels, dt) fills an open-addressing table of `capacity`
 *     slots (a power of two, at least twice the number of distinct labels): keys_dev[s] = label
 *     (0 = empty slot; label 0 is background and skipped), count_dev[s] = its voxels,
 *     max_dev[s] = its largest distance, argmax_dev[s] = the smallest linear index (x fastest)
 *     where that maximum is attained, box_dev[6*s..] = bounding box x0 y0 z0 x1 y1 z1 (inclusive).
 *     *overflow_dev is raised when the table was too small (repeat with a larger capacity).
 * edtb200_label_extract : out = (labels == key) ? dt : 0 inside `box` (6 host ints as above, NULL =
 *     whole volume), nothing outside it is touched -- the equivalent of transfer_run_voxels on a
 *     blank image; erase != 0 zeroes the box instead (the reference's erase()).
 * All pointers except `box` are device pointers on `device`; asynchronous on `stream`. */
int edtb200_label_stats(const void *labels_dev, int label_bytes, const float *dt_dev,
                        int64_t sx, int64_t sy, int64_t sz, int capacity,
                        unsigned long long *keys_dev, unsigned long long *count_dev, float *max_dev,
                        long long *argmax_dev, int *box_dev, int *overflow_dev, int device, void *stream);

int edtb200_label_extract(const void *labels_dev, int label_bytes, const float *dt_dev,
                          int64_t sx, int64_t sy, int64_t sz, unsigned long long key, const int *box,
                          int erase, float *out_dev, int device, void *stream);

/* Page-locked host memory for result arrays (cudaHostAlloc, portable): a transform whose `out` lies
 * in such a block is copied back by one DMA at PCIe rate, with no staging copy and no page faults.
 * The Python front door keeps a small pool of these blocks behind the arrays it returns.
 * edtb200_host_alloc returns NULL when no device / no memory is available. */
void *edtb200_host_alloc(size_t bytes);
void edtb200_host_free(void *p);

/* Measurement hooks (used by bench.py): with profiling enabled on the calling thread, every
 * edtb200_transform records CUDA events around its axis passes on the transform's stream into a
 * ring of 256 slots -- nothing synchronises inside a timed loop.  edtb200_pass_ms(k, ms) waits
 * for the transform issued k calls ago (0 = the latest) and returns the device time of its
 * first, second and third axis pass in milliseconds (0 for passes that did not run). */
int edtb200_profile_passes(int enable);
int edtb200_pass_ms(int steps_back, float *ms3);

/* Free every cached device buffer / stream this library holds on all devices. */
int edtb200_release(void);

#ifdef __cplusplus
}
#endif
#endif /* EDT_B200_H */
