/*
 * gtb200.h -- C ABI of libgtb200.so, the B200 (sm_100a) stencil execution + halo-exchange library that sits
 * behind the GridTools backend-tag / gcl interfaces.
 *
 * Every entry point replaces a piece of the reference's `stencil::gpu` / `gcl::gpu` path; the reference
 * interface each one stands in for is cited as file:line (paths relative to the GridTools v2.4.0 tree).
 *
 * Conventions
 *  - Plain C: pointers, sizes, ints.  No C++/torch types.  All functions return a gtb_status (0 == OK) and
 *    never throw; gtb_last_error() gives the message of the last failure on the calling thread
 *    (the reference throws std::runtime_error from GT_CUDA_CHECK, common/cuda_util.hpp:20-35 -- the C++
 *    header include/gtb200/stencil/b200.hpp turns a non-zero status back into that exception).
 *  - A field is {pointer to element (0,0,0) of the COMPUTE DOMAIN, element strides}.  This is exactly what a
 *    backend receives from the frontend: origin-shifted SIDs (stencil/core/backend.hpp:29-34) whose
 *    sid::get_origin / sid::get_strides give pointer and strides.  Halo points are addressed with negative or
 *    >= n indices and must be valid device memory within the stencil's extent (frontend/run.hpp:215-232).
 *  - Fields are borrowed for the duration of the call (frontend/run.hpp:214).  Kernels are enqueued on `stream`
 *    (a cudaStream_t passed as void*, NULL = legacy default stream like the reference,
 *    stencil/gpu/launch_kernel.hpp:161) and the call returns without synchronising, like
 *    common/cuda_util.hpp:79-96 in release mode.
 *  - There is no CPU fallback anywhere in this library: without a CUDA device every compute entry point
 *    returns GTB_ERR_CUDA.
 */
#ifndef GTB200_H
#define GTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTB_VERSION 100 /* 0.1.0 */

typedef enum {
    GTB_OK = 0,
    GTB_ERR_ARG = 1,    /* null pointer, negative size, unsupported element size */
    GTB_ERR_LAYOUT = 2, /* layout the kernel cannot address (stride_i != 1 for the tiled kernels) */
    GTB_ERR_CUDA = 3,   /* CUDA runtime / driver error, see gtb_last_error() */
    GTB_ERR_ALLOC = 4,  /* scratch allocation failed */
    GTB_ERR_STATE = 5   /* halo object used out of order (e.g. exchange before connect) */
} gtb_status;

typedef struct {
    void *ptr;                            /* element (0,0,0) of the compute domain (device memory) */
    int64_t stride_i, stride_j, stride_k; /* element strides */
} gtb_field;

/* ------------------------------------------------------------------------------------------- runtime */

/* Library version (GTB_VERSION). */
int gtb_version(void);
/* Message of the last error on this thread ("" if none).  Replaces the what() of the std::runtime_error thrown by
 * GT_CUDA_CHECK (common/cuda_util.hpp:20-35). */
const char *gtb_last_error(void);
/* Number of visible CUDA devices (0 if none / no driver).  Never fails. */
int gtb_device_count(void);
/* Selects `device` for the calling thread and warms the kernels' function attributes (the reference calls
 * cudaFuncSetAttribute on every launch, common/cuda_util.hpp:84-88; here it happens once). */
int gtb_init(int device);
/* SM count and L2 size of the current device (used by the host side to size persistent grids / report). */
int gtb_device_info(int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes);
/* Tuning knob: kernels pick a variant from integer options, e.g. ("hd.variant", 0=auto 1=cp.async 2=tma),
 * ("va.threads", 32..256), ("va.unroll", 1..8), ("va.scratch", 0=auto 1=global 2=smem).  Unknown keys return
 * GTB_ERR_ARG.  The reference's only knobs are the compile-time block sizes of gpu<IBlock,JBlock,KBlock>
 * (stencil/gpu/entry_point.hpp:147-150). */
int gtb_set_option(const char *key, int value);
int gtb_get_option(const char *key, int *value);
/* Releases the cached scratch (temporaries) of the calling device; the reference keeps them in a thread-local
 * sid::device::cached_allocator (sid/allocator.hpp:65-95, stencil/gpu/entry_point.hpp:155-175). */
int gtb_release_scratch(void);
/* Diagnosis only: with option ("va.debug", 128) the paired-warp vertical advection kernel writes globaltimer stamps
 * per warp pair ([pair][32] int64: [0] start, forward pass p [1+4p] begin / [2+4p] end, backward pass p [3+4p] / [4+4p]);
 * this synchronises the device, copies up to `bytes` of them to `dst` (host) and clears them. */
int gtb_debug_trace(void *dst, int64_t bytes);
/* Host <-> device copy of the ni x nj x nk sub-box that starts at the given origins (same element strides on both sides,
 * i contiguous) on `stream`: storage::gpu's update_target / update_host (storage/gpu.hpp:86-99, one blocking cudaMemcpy of
 * the whole padded allocation) restricted to the points a stencil reads or writes.  Asynchronous when the host memory is
 * pinned.  to_device != 0: host -> device. */
int gtb_copy_box_async(void *device_origin, void *host_origin, int elem_size, int64_t stride_j, int64_t stride_k, int ni,
    int nj, int nk, int to_device, void *stream);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t gtb_launch_count(void);
/* Name of the kernel the calling thread launched last ("" before the first): which variant the automatic choice took. */
const char *gtb_last_kernel(void);

/* --------------------------------------------------------------------------------- named stencil kernels
 * Each is the fused B200 kernel for one spec of the reference's regression/perf suite; together they are what
 * `run(spec, gpu<>{}, grid, fields...)` (frontend/run.hpp:242-250 -> stencil/gpu/entry_point.hpp:254-260)
 * executes for that spec.  ni,nj,nk = compute-domain size (grid.i_size(), j_size(), k_size()). */

/* copy_stencil.cpp:24-36 : out = in (bit exact).  elem_size in {4,8}.  Any strides. */
int gtb_copy(const gtb_field *in, const gtb_field *out, int ni, int nj, int nk, int elem_size, void *stream);

/* horizontal_diffusion.cpp:35-106 : execute_parallel, ij_cached(lap, flx, fly), 4 stages fused into one kernel.
 * Reads `in` on [-2, n+2) in i and j.  stride_i must be 1.  `out` must not alias `in`/`coeff`. */
int gtb_hori_diff_f64(const gtb_field *in, const gtb_field *coeff, const gtb_field *out, int ni, int nj, int nk,
    void *stream);
int gtb_hori_diff_f32(const gtb_field *in, const gtb_field *coeff, const gtb_field *out, int ni, int nj, int nk,
    void *stream);

/* simple_hori_diff.cpp:25-88 : execute_parallel, ij_cached(lap), wlap_function + divflux_function in one kernel.
 * crlato / crlatu are the reference's j-only fields (builder selector<0,1,0>): {pointer to the compute domain's j = 0,
 * stride_j}, read on j in [-1, nj]; stride_i / stride_k are ignored.  Reads `in` on [-2, n+2) in i and j. */
int gtb_simple_hori_diff_f64(const gtb_field *in, const gtb_field *coeff, const gtb_field *crlato,
    const gtb_field *crlatu, const gtb_field *out, int ni, int nj, int nk, void *stream);
int gtb_simple_hori_diff_f32(const gtb_field *in, const gtb_field *coeff, const gtb_field *crlato,
    const gtb_field *crlatu, const gtb_field *out, int ni, int nj, int nk, void *stream);

/* vertical_advection_dycore.cpp:32-149 : forward sweep (k_cached ccol/dcol flush, u_stage fill) + backward sweep
 * (k_cached data_col) fused into one kernel; utens_stage is updated in place.  Reads wcon at i+1 and k+1.
 * nk >= 2.  stride_i must be 1. */
int gtb_vert_adv_f64(const gtb_field *utens_stage, const gtb_field *u_stage, const gtb_field *wcon,
    const gtb_field *u_pos, const gtb_field *utens, double dtr_stage, int ni, int nj, int nk, void *stream);
int gtb_vert_adv_f32(const gtb_field *utens_stage, const gtb_field *u_stage, const gtb_field *wcon,
    const gtb_field *u_pos, const gtb_field *utens, float dtr_stage, int ni, int nj, int nk, void *stream);

/* tridiagonal.cpp:39-97 : Thomas solve, forward + backward in one kernel.  sup and rhs are overwritten by the
 * forward sweep exactly like in the reference (they are inout fields there). */
int gtb_tridiagonal_f64(const gtb_field *inf, const gtb_field *diag, const gtb_field *sup, const gtb_field *rhs,
    const gtb_field *out, int ni, int nj, int nk, void *stream);

/* advection_pdbott_prepare_tracers.cpp:23-34 run through expandable_run<2> (frontend/expandable_run.hpp:144-184):
 * out[t] = rho * in[t] for all n_tracers in ONE launch (the reference needs ceil(n/2) launches). */
int gtb_prepare_tracers_f64(const gtb_field *out, const gtb_field *in, int n_tracers, const gtb_field *rho, int ni,
    int nj, int nk, void *stream);

/* ------------------------------------------------------------------------------------------ gcl halo exchange
 * Replaces hndlr_dynamic_ut<..., gpu> (gcl/high_level/descriptors_manual_gpu.hpp:83-515), the 24 pack/unpack
 * kernels gcl/high_level/m_{pack,unpack}{X,Y,Z}{L,U}.hpp and the MPI choreography of
 * gcl/low_level/Halo_Exchange_3D.hpp:145-931 for ranks that live on one NVLink/NVSwitch box.
 *
 * One gtb_halo object per rank (process, one GPU each).  Dimension d of a descriptor is the d-th STORAGE
 * dimension in increasing-stride order (d = 0 is the unit-stride axis); halo descriptors have the meaning of
 * common/halo_descriptor.hpp:44-227 (minus, plus, begin, end inclusive, total length).
 * Neighbours are indexed n = (e0+1) + 3*(e1+1) + 9*(e2+1) with e_d in {-1,0,1} the offset along storage
 * dimension d; n = 13 is the rank itself (unused).
 *
 * Life cycle (mirrors halo_exchange_dynamic_ut, gcl/halo_exchange.hpp:163-306):
 *   gtb_halo_create            <- ctor + add_halo<D>() x3 + setup(max_fields)            (:202,:235,:216)
 *   gtb_halo_export/connect    <- buffer registration of Halo_Exchange_3D (:270-287 of descriptors_manual_gpu.hpp);
 *                                 the blobs travel through whatever out-of-band channel the host has
 *                                 (torch.distributed, MPI, a file)
 *   gtb_halo_pack              <- pack(vector<T*>)                                        (:250,:269)
 *   gtb_halo_exchange          <- exchange() = start_exchange() + wait()                  (:284-304)
 *   gtb_halo_unpack            <- unpack(vector<T*>)                                      (:260,:276)
 *   gtb_halo_destroy           <- dtor
 */
typedef struct {
    int minus, plus, begin, end, total;
} gtb_halo_desc;

typedef struct gtb_halo gtb_halo;

#define GTB_HALO_BLOB_BYTES 512

/* neighbour_rank[27]: rank of the neighbour in direction n or -1 (non-periodic border,
 * gcl/low_level/proc_grids_3D.hpp:179-211); entry 13 is ignored.  The handle owns 26 send + 26 recv device buffers
 * sized like descriptors_manual_gpu.hpp:255-300: prod_d s_length(e_d) * max_fields * elem_size. */
int gtb_halo_create(const gtb_halo_desc desc[3], const int neighbour_rank[27], int my_rank, int max_fields,
    int elem_size, gtb_halo **out);
int gtb_halo_destroy(gtb_halo *h);

/* Bytes of the message towards / from neighbour n for n_fields fields (0 if there is no such neighbour). */
int64_t gtb_halo_send_bytes(const gtb_halo *h, int n, int n_fields);
int64_t gtb_halo_recv_bytes(const gtb_halo *h, int n, int n_fields);
/* Device pointers of the staging buffers, for hosts that move the messages themselves (NCCL / MPI transport). */
void *gtb_halo_send_buffer(const gtb_halo *h, int n);
void *gtb_halo_recv_buffer(const gtb_halo *h, int n);

/* Peer-to-peer transport over NVLink: export this rank's receive arena (cudaIpcMemHandle + layout) as an opaque
 * blob of GTB_HALO_BLOB_BYTES, hand every neighbour's blob to connect (blobs[n] = blob of neighbour_rank[n], NULL
 * where there is none; a neighbour that is this very process -- periodic grid of extent 1 or 2 ranks in one
 * process -- is detected and not re-opened). */
int gtb_halo_export(gtb_halo *h, void *blob);
int gtb_halo_connect(gtb_halo *h, const void *const blobs[27]);

/* One fused launch: gathers the send regions of all fields for all existing neighbours into the send buffers.
 * fields[f] points at storage element (0,0,0) INCLUDING the halo (like the raw T* the reference takes). */
int gtb_halo_pack(gtb_halo *h, void *const *fields, int n_fields, void *stream);
/* Fused pack + NVLink store + signal in ONE launch: writes every message straight into the neighbour's receive
 * buffer; the last block of each direction raises the neighbour's arrival flag (needs connect).  Equivalent to
 * pack() + the do_sends() half of exchange(). */
int gtb_halo_pack_send(gtb_halo *h, void *const *fields, int n_fields, void *stream);
/* Pushes the packed send buffers into the neighbours' receive buffers and raises their flags (needs connect). */
int gtb_halo_send(gtb_halo *h, int n_fields, void *stream);
/* Device-side wait until all expected messages of the current epoch have arrived (enqueued on stream). */
int gtb_halo_wait(gtb_halo *h, void *stream);
/* One fused launch: scatters every received message into the halo regions of all fields. */
int gtb_halo_unpack(gtb_halo *h, void *const *fields, int n_fields, void *stream);
/* wait + unpack in ONE launch: every block acquires the arrival flag of its direction before it scatters that
 * message (needs connect).  Together with gtb_halo_pack_send an exchange is two launches and no host synchronisation,
 * against up to 12 x n_fields launches, a cudaDeviceSynchronize and 2 x 26 MPI calls in the reference. */
int gtb_halo_wait_unpack(gtb_halo *h, void *const *fields, int n_fields, void *stream);
/* pack_send + wait_unpack + next_epoch in one call (two launches on `stream`, or one with option "halo.fused"): the whole
 * pack() / exchange() / unpack() sequence of halo_exchange_dynamic_ut (gcl/halo_exchange.hpp:250-304) for hosts that do
 * not need the phases separately. */
int gtb_halo_exchange(gtb_halo *h, void *const *fields, int n_fields, void *stream);
/* halo_exchange_generic (gcl/halo_exchange.hpp:335-513, high_level/descriptor_generic_manual.hpp:370-796): every field
 * brings ITS OWN halo descriptors (field_on_the_fly, high_level/field_on_the_fly.hpp:27-95).  The handle is created
 * from the "halo example" (gtb_halo_create with the enclosing descriptors, max_fields and the word size: buffers hold
 * max_fields fields of the example's regions, like hndlr_generic::setup :389-393).  All fields of a call travel in ONE
 * message per neighbour (fields concatenated in argument order), packed by one launch and unpacked by one launch.
 * desc[] is in increasing-stride order and in units of the handle's elem_size words (an element of k words scales
 * the descriptor of dimension 0 by k).  GTB_ERR_ARG if a message would not fit the buffers. */
typedef struct {
    void *ptr; /* storage element (0,0,0), halo included */
    gtb_halo_desc desc[3];
} gtb_halo_field;
int gtb_halo_generic_pack_send(gtb_halo *h, const gtb_halo_field *fields, int n_fields, void *stream);
int gtb_halo_generic_wait_unpack(gtb_halo *h, const gtb_halo_field *fields, int n_fields, void *stream);
/* Device-side waits for a neighbour's message give up after option "halo.timeout_ms" (default 60 000; 0 = wait for
 * ever, like the MPI_Wait of Halo_Exchange_3D).  A wait that gave up does NOT unpack that message (the halo keeps its
 * old values) and stores 1 + direction in an error word in mapped host memory; every later wait of the object fails
 * at once.  gtb_halo_error synchronises the device first, gtb_halo_poll_error just reads the word (cheap enough for
 * every time step); *code = 0 if all is well. */
int gtb_halo_error(gtb_halo *h, int *code);
int gtb_halo_poll_error(gtb_halo *h, int *code);
/* Diagnosis: with a device buffer of 256 x 8 uint64 set, the transfer kernels of epoch e stamp %globaltimer (ns) into
 * row e % 256: [0] pack starts, [1] pack ends (flags raised), [2] unpack starts, [3] last arrival flag acquired,
 * [4] unpack ends.  NULL switches it off.  gtb_stamp enqueues a one-thread kernel that stores %globaltimer. */
int gtb_halo_set_trace(gtb_halo *h, void *device_u64_256x8);
int gtb_stamp(void *device_u64, void *stream);
/* Advances the epoch after unpack (double-buffered arenas: a neighbour may already send epoch e+1 while this rank
 * still unpacks epoch e). */
int gtb_halo_next_epoch(gtb_halo *h);

/* ------------------------------------------------------------------------------------------ boundary conditions
 * boundaries/boundary.hpp:57-72 (`boundary<BoundaryFunction, Arch, Predicate>::apply`) for the predefined conditions:
 * GTB_BC_VALUE = value_boundary<T> / zero_boundary (value.hpp:27-66, zero.hpp: every field is set to `value`),
 * GTB_BC_COPY = copy_boundary (copy.hpp:26-46: fields[0 .. n-2] receive fields[n-1]).  The condition is applied on
 * the OUTSIDE region (halo_descriptor::loop_{low,high}_bound_outside, apply.hpp:44-56) of every direction n with
 * direction_mask[n] != 0 (the Predicate; NULL = default_predicate = all 26).  Fields as in the gcl calls: pointer to
 * storage element (0,0,0) including the halo, layout given by the descriptors' total lengths.  One launch for all
 * directions and fields (apply_gpu.hpp:236-313 is one launch per call too, but per-direction 3-d thread blocks). */
typedef enum { GTB_BC_VALUE = 0, GTB_BC_COPY = 1 } gtb_bc_kind;
int gtb_boundary_apply(const gtb_halo_desc desc[3], const int direction_mask[27], int kind, double value,
    void *const *fields, int n_fields, int elem_size, void *stream);
/* distributed_boundaries.hpp:141-200 (exchange, then the boundary condition where the process grid has no neighbour,
 * proc_grid_predicate): after this call every unpack / wait_unpack / exchange launch of the halo object also writes
 * `value` into the outside regions of the directions without a neighbour -- the condition costs no extra launch.
 * kind = GTB_BC_VALUE, or -1 to switch it off again. */
int gtb_halo_set_boundary(gtb_halo *h, int kind, double value);

/* ------------------------------------------------------------------------------------------ device-side gates
 * A halo exchange that runs on its own stream beside the stencils has to be ordered against them twice per step: the
 * stencil must not read a halo before it is unpacked, and the unpack of a later exchange must not overwrite a halo a
 * stencil launch is still reading.  Stream events do that (and the reference's host-synchronous exchange does it by
 * blocking), but an event record plus a cross-stream wait around every launch cost ~5 us of a 25-60 us step.  These
 * calls move both orderings onto the device:
 *   gtb_halo_unpacked_flag(h)         device address of a uint64 that holds the epoch (1, 2, ...: one per exchange,
 *                                     gtb_halo_epoch(h) is the epoch the NEXT exchange will carry) of the last
 *                                     completed wait_unpack of this handle;
 *   gtb_stencil_gate(flag, v, post)   one-shot, for the next gtb_hori_diff_* / gtb_vert_adv_f64 launch of the calling
 *                                     thread: the kernel waits on the device until *flag >= v before it reads global
 *                                     memory (flag = NULL: no wait) and adds 1 to the uint64 *post when all of it is
 *                                     done (post = NULL: nothing).  Needs option "reserve_sms" >= 1: the kernel that
 *                                     raises the flag must find SMs the spinning stencil does not occupy
 *                                     (GTB_ERR_STATE otherwise).  Only the default kernels take a gate (GTB_ERR_ARG);
 *   gtb_halo_gate(h, counter, v)      one-shot, for the next wait_unpack / exchange of h: the unpack waits on the
 *                                     device until *counter >= v (typically `post` of the stencil launch that last
 *                                     read these halos) before it scatters.
 * A wait gives up after ~0.2 s instead of hanging the device; gtb_gate_timeouts() synchronises the device and reports
 * how many waits did so far (0 in a correct program). */
void *gtb_halo_unpacked_flag(gtb_halo *h);
uint64_t gtb_halo_epoch(const gtb_halo *h);
int gtb_stencil_gate(const void *wait_flag, uint64_t wait_value, void *post_counter);
int gtb_halo_gate(gtb_halo *h, const void *counter, uint64_t value);
int gtb_gate_timeouts(int64_t *count);

/* ------------------------------------------------------------------------------------------ tensor maps
 * A TMA descriptor (CUtensorMap, 128 bytes, 64-byte aligned) of a 3-d box over an i-contiguous field, for kernels that
 * are compiled in the user's translation unit (the generic fused path of include/gtb200/stencil/b200_fused.hpp stages
 * read-only fields through shared memory with it).  dims: extents of the tensor from `base`; strides_bytes: j and k;
 * box: elements per dimension.  GTB_ERR_LAYOUT when the field is not TMA-addressable. */
int gtb_tensor_map_3d(void *map128, const void *base, int elem_size, const int64_t dims[3],
    const int64_t strides_bytes[2], const int box[3]);

/* ------------------------------------------------------------------------------------------ staged copies
 * Whole-allocation transfers between PAGEABLE host memory and the device for storage traits (the host mirror of a
 * GridTools data_store is a new[] array the traits cannot pin, storage/data_store.hpp:101-104): chunks are staged by
 * several host threads through a ring of pinned buffers and overlap with the asynchronous copies, instead of the single
 * blocking cudaMemcpy of storage/gpu.hpp:86-99.  Stream-ordered on `stream` (NULL = legacy default stream).  Upload:
 * returns when host_src may be reused; download: returns when host_dst is complete. */
/* device memory for host bindings that do not include the CUDA runtime (storage traits) */
int gtb_device_malloc(void **out, int64_t bytes);
int gtb_device_free(void *p);
/* Page-locked host memory: a host mirror allocated with it is copied by the copy engine directly (gtb_staged_* skip the
 * staging ring for pinned memory, whoever pinned it).  The reference's data_store allocates its mirror itself
 * (storage/data_store.hpp:101-104); patches/gridtools-host-mirror-through-traits.patch routes that through the traits. */
int gtb_host_malloc(void **out, int64_t bytes);
int gtb_host_free(void *p);
int gtb_staged_upload(void *device_dst, const void *host_src, int64_t bytes, void *stream);
int gtb_staged_download(void *host_dst, const void *device_src, int64_t bytes, void *stream);

/* ------------------------------------------------------------------------------------------------- streams
 * For host bindings that do not include the CUDA runtime themselves.  gtb_stream_create returns a NON-BLOCKING
 * cudaStream_t (optionally of the highest priority, for exchanges that run beside a stencil); gtb_stream_after_default
 * makes the work issued to it from now on wait for everything issued to the legacy default stream so far;
 * gtb_stream_synchronize blocks the host until the stream is idle. */
int gtb_stream_create(void **stream, int high_priority);
int gtb_stream_destroy(void *stream);
int gtb_stream_after_default(void *stream);
int gtb_stream_synchronize(void *stream);

/* ------------------------------------------------------------------------------------- recorded call sequences
 * The reference's user programs drive their time loop from C++ (tests/regression/gcl/copy_stencil_parallel.cpp:126-145:
 * he.pack / he.exchange / he.unpack followed by run(spec, backend, grid, fields...)), a microsecond or two of host
 * time per call.  Hosts that reach this library through ctypes / JNI / cgo pay several microseconds per call, more
 * than a 256x256x80 stencil step leaves.  A gtb_seq records such a loop once -- stencil launches, halo exchanges and
 * the event record / wait operations that order a compute stream against a communication stream -- and
 * gtb_seq_run() issues any slice of it in recorded order with ONE call.  Fields and halo objects are borrowed and must
 * outlive the sequence.  Event numbers are small non-negative slots owned by the sequence (timing disabled);
 * a wait refers to the most recent record of that slot at the time it is issued, as with cudaStreamWaitEvent. */
typedef struct gtb_seq gtb_seq;
int gtb_seq_create(gtb_seq **out);
int gtb_seq_destroy(gtb_seq *s);
int gtb_seq_size(const gtb_seq *s);
/* elem_size 8 -> gtb_hori_diff_f64 / gtb_vert_adv_f64, 4 -> the _f32 entry points */
int gtb_seq_add_hori_diff(gtb_seq *s, int elem_size, const gtb_field *in, const gtb_field *coeff, const gtb_field *out,
    int ni, int nj, int nk, void *stream);
int gtb_seq_add_vert_adv(gtb_seq *s, int elem_size, const gtb_field *utens_stage, const gtb_field *u_stage,
    const gtb_field *wcon, const gtb_field *u_pos, const gtb_field *utens, double dtr_stage, int ni, int nj, int nk,
    void *stream);
int gtb_seq_add_prepare_tracers(gtb_seq *s, const gtb_field *out, const gtb_field *in, int n_tracers,
    const gtb_field *rho, int ni, int nj, int nk, void *stream);
int gtb_seq_add_halo_exchange(gtb_seq *s, gtb_halo *h, void *const *fields, int n_fields, void *stream);
/* gtb_stencil_gate / gtb_halo_gate for the operation recorded NEXT (armed when the sequence reaches them) */
int gtb_seq_add_stencil_gate(gtb_seq *s, const void *wait_flag, uint64_t wait_value, void *post_counter);
int gtb_seq_add_halo_gate(gtb_seq *s, gtb_halo *h, const void *counter, uint64_t value);
int gtb_seq_add_record(gtb_seq *s, int event, void *stream);
int gtb_seq_add_wait(gtb_seq *s, void *stream, int event);
/* Timing marks: a mark is a timing-enabled event recorded in `stream` when the sequence reaches it, so that a region
 * in the MIDDLE of one gtb_seq_run() slice can be timed on the device (ranks of a multi-GPU loop are in lock-step
 * there, which they are not at the first operation after a host barrier).  gtb_seq_elapsed_ms waits for mark_b. */
int gtb_seq_add_mark(gtb_seq *s, int mark, void *stream);
int gtb_seq_elapsed_ms(gtb_seq *s, int mark_a, int mark_b, float *ms);
/* diagnosis: gtb_stamp(device_u64, stream) when the sequence reaches this point */
int gtb_seq_add_stamp(gtb_seq *s, void *device_u64, void *stream);
/* Issues operations [first, first + count) of the sequence. */
int gtb_seq_run(gtb_seq *s, int first, int count);

#ifdef __cplusplus
}
#endif
#endif /* GTB200_H */
