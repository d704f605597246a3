/*
 * dmk.h -- C ABI of libdmk.so: B200 (sm_100a) channel-generation kernels for DeepMIMO.
 *
 * Drop-in boundary.  The reference (jmoraispk/DeepMIMO v4.0.0a3) has no FFI: its seam for
 * this path is the Python method
 *     Dataset.compute_channels(params) -> complex64 [n_ue, M_rx, M_tx, K | P]
 *         deepmimo/generator/dataset.py:224-268
 * which runs rotation (dataset.py:310-356, geometry.py:244-319), FoV (dataset.py:461-512,
 * geometry.py:162-195), element patterns (dataset.py:665-691, ant_patterns.py:21-71),
 * array responses (dataset.py:380-417, geometry.py:38-120), per-path OFDM gains
 * (channel.py:170-198) and the per-user accumulation (channel.py:200-289).  The entry points
 * below are what a binding for that method calls: one launch computes all of the above for a
 * contiguous range of users.  deepmimo_b200/channels.py is the ctypes binding; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - All array pointers are DEVICE pointers owned by the caller (e.g. torch tensors), except
 *    `const dmk_desc*` which is a HOST pointer read during the call.  The library allocates
 *    nothing persistent, keeps no global state besides a thread-local error string and a
 *    launch counter, and is stream-ordered on `cuda_stream` (a cudaStream_t, NULL = legacy
 *    default stream).
 *  - Path matrices are float32 row-major [n, n_cols] with row stride `ld` (elements), valid
 *    paths anywhere in the row, NaN = no path (channel.py:260).
 *  - Return value: 0 on success, negative dmk_status on error (dmk_last_error() has the text).
 */
#ifndef DMK_H
#define DMK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMK_ABI_VERSION 3
#define DMK_MAX_PATHS   32    /* path columns per user handled by one launch (reference: MAX_PATHS = 25, consts.py:180) */
#define DMK_MAX_TIMES   4096  /* time snapshots per launch */

typedef enum dmk_status {
    DMK_OK                =  0,
    DMK_ERR_INVALID_ARG   = -1,
    DMK_ERR_UNSUPPORTED   = -2,   /* unknown pattern (ant_patterns.py:119-122); rx_filter=1 with more than 9728 subcarriers */
    DMK_ERR_CUDA          = -3
} dmk_status;

typedef enum dmk_pattern {
    DMK_PATTERN_ISOTROPIC       = 0,  /* ant_patterns.py:21-31  */
    DMK_PATTERN_HALFWAVE_DIPOLE = 1   /* ant_patterns.py:34-71  */
} dmk_pattern;

/* Channel-generation descriptor: the numeric content of ChannelGenParameters
 * (deepmimo/generator/channel.py:33-63) plus the FoV set by Dataset.apply_fov
 * (deepmimo/generator/dataset.py:423-448). */
typedef struct dmk_desc {
    int32_t bs_shape[2];        /* bs_antenna.shape[0:2]: elements along y, z (geometry.py:105-120) */
    int32_t ue_shape[2];        /* ue_antenna.shape[0:2]                                            */
    double  bs_spacing;         /* wavelengths (kd = 2*pi*spacing, dataset.py:393)                  */
    double  ue_spacing;
    double  bs_rot_deg[3];      /* rotation about x, y, z in degrees (channel.py:38)                */
    double  ue_rot_deg[3];      /* used for every user when the ue_rot_deg argument is NULL         */
    double  bs_fov_deg[2];      /* [horizontal, vertical] degrees                                   */
    double  ue_fov_deg[2];
    int32_t fov_side_enabled[2];/* [bs, ue]: 1 = side is restricted (not _is_full_fov, dataset.py:450-459) */
    int32_t fov_any;            /* 1 = the reference builds a mask (dataset.py:493-512); 0 = mask is None */
    int32_t pattern[2];         /* [bs, ue] dmk_pattern                                             */
    int32_t num_paths;          /* params.num_paths: first num_paths columns are used (dataset.py:255-262) */
    int32_t n_cols;             /* columns present in the path matrices (<= DMK_MAX_PATHS)          */
    int32_t n_subcarriers;      /* ofdm.subcarriers (N)                                             */
    int32_t n_selected;         /* len(ofdm.selected_subcarriers) (K); ignored by dmk_channels_td   */
    const int32_t *subcarriers; /* DEVICE pointer, [K] selected subcarrier indices                  */
    int32_t subc_start;         /* if subc_step != 0 the selection is start + step*i and            */
    int32_t subc_step;          /*   `subcarriers` may be NULL                                      */
    double  bandwidth;          /* Hz (Ts = 1/bandwidth, channel.py:223)                            */
    int32_t rx_filter;          /* 0 | 1: receive low-pass filter of the OFDM branch (channel.py:166-168, :193-194); ignored by dmk_channels_td */
    int32_t n_times;            /* 0 = no trailing time axis; T >= 1 appends [.., T] (row a11)      */
    const double *times;        /* DEVICE pointer, [T] snapshot times in seconds                    */
    int32_t flags;              /* DMK_FLAG_* (ABI 2); 0 = plain stream-ordered launch                   */
    int32_t kernel_hint;        /* dmk_kernel_hint (ABI 3); 0 = the library picks the kernel by shape     */
    int32_t ws_helpers;         /* 0 = by shape; 1, 2 or 4 pins the helper-warp count of the persistent tensor-core kernel; with
                                 * the warp-level-kernel hint (7): 16 or 32 pins the chunk width of the warp-level kernel (tests, A/B timing) */
    int32_t ws_split;           /* 0 = by shape; >= 1 pins how many work items a user's stages are dealt into; with
                                 * the warp-level-kernel hint (7): m-tiles resident per group (tests, A/B timing) */
} dmk_desc;

/* dmk_desc.kernel_hint: force a kernel family where the shape is eligible for it (parity tests run every family on the same
 * inputs; A/B timing).  A hint the shape is not eligible for falls through to the next family, exactly like the automatic choice. */
typedef enum dmk_kernel_hint {
    DMK_KERNEL_AUTO  = 0,
    DMK_KERNEL_TILE  = 1,   /* generic tile kernel (any subcarrier list, time axis, rx_filter)      */
    DMK_KERNEL_FFMA  = 2,   /* packed-FP32 CUDA-core kernel                                         */
    DMK_KERNEL_TC    = 3,   /* persistent warp-specialised tcgen05 kernel                           */
    DMK_KERNEL_TC1   = 4,   /* one-CTA-per-user tcgen05 kernel                                      */
    DMK_KERNEL_SMALL = 5,   /* small-array kernel (M <= 16), densely packed                         */
    DMK_KERNEL_SMALL1 = 6,  /* round-1 small-array kernel (one warp per user), kept for A/B timing  */
    DMK_KERNEL_MMA   = 7,   /* warp-level tensor-core kernel for small per-user outputs (M <= 1024, K <= 4096) */
    DMK_KERNEL_ROWS  = 8    /* warp-per-user kernel with lanes = antenna rows for K <= 8 selected subcarriers  */
} dmk_kernel_hint;

/* dmk_desc.flags.
 * DMK_FLAG_INDEPENDENT_LAUNCH: the caller asserts that this launch neither reads nor overwrites anything the
 * kernel launched immediately before it on the same stream writes (e.g. consecutive user chunks of one
 * compute_channels call going to different output buffers).  The persistent tensor-core kernel then starts
 * filling SMs while the previous launch drains its tail (programmatic dependent launch without a grid
 * dependency wait).  Without the flag every launch observes full stream order.
 * Ordering contract: a flagged launch begins only after every CTA of the previous launch has begun, and an
 * unflagged launch begins only after everything before it has completed.  A flagged launch that follows s - 1
 * other flagged launches can therefore overlap with its s predecessors but never with anything older than the
 * last unflagged launch's predecessors: with a ring of R output buffers, at most R - 1 consecutive launches may
 * carry the flag before one launch goes without it (deepmimo_b200.iter_channels does exactly that). */
#define DMK_FLAG_INDEPENDENT_LAUNCH 1
/* DMK_FLAG_F64_INPUTS (ABI 3): the seven path-matrix arguments of every entry point address float64 arrays (passed through the
 * `const float *` parameters; `ld` counts float64 elements).  The prologue then follows NumPy's all-float64 flow for float64
 * inputs: deg2rad, sin/cos, 10**(p/10), toa/Ts and the path gain in float64.  `doppler_hz` stays float32. */
#define DMK_FLAG_F64_INPUTS 2

/* Frequency-domain channels (freq_domain = 1):
 *   out[u, r, t, k (, it)] = sum_p c_p a_rx[r,p] a_tx[t,p] exp(-j 2 pi k delay_n[p] / N) (* exp(+j 2 pi f_D[p] t_it))
 * (rx_filter = 1: the delay phasor is replaced by sum_d sinc(d - delay_n[p]) exp(-j 2 pi d k / N), d = 0..N-1)
 * complex64, C-contiguous [n, M_r, M_t, K (, T)], every element written (users without paths -> zeros).
 *   ue_rot_deg : NULL or DEVICE double [n,3] per-user UE rotation (dataset.py:328-338)
 *   doppler_hz : NULL or DEVICE float  [n, n_cols] (ld) per-path Doppler shift in Hz (row a11 extension)
 *   fov_mask   : NULL or DEVICE uint8 [n, n_cols]  = Dataset._fov_mask (all ones when fov_any == 0)
 *   valid_mask : NULL or DEVICE uint8 [n, n_cols]  = ~isnan(power) for the first num_paths columns (channel.py:260)
 *   clip_mask  : NULL or DEVICE uint8 [n, n_cols]  = valid & (delay_n >= N) (channel.py:187)
 */
int dmk_channels_fd(const dmk_desc *desc,
                    const float *power_dbw, const float *phase_deg, const float *delay_s,
                    const float *aoa_az_deg, const float *aoa_el_deg,
                    const float *aod_az_deg, const float *aod_el_deg,
                    const double *ue_rot_deg, const float *doppler_hz,
                    int64_t n_users, int32_t ld,
                    void *out_c64,
                    uint8_t *fov_mask, uint8_t *valid_mask, uint8_t *clip_mask,
                    void *cuda_stream);

/* Time-domain channels (freq_domain = 0):
 *   out[u, r, t, j (, it)] = a_rx[r,p_j] a_tx[t,p_j] sqrt(power[p_j]) exp(j phase[p_j]) (* exp(+j 2 pi f_D t_it)),
 * p_j = j-th valid path of user u; slots j >= n_valid are zero (channel.py:285-287).
 * complex64 [n, M_r, M_t, P (, T)], P = min(num_paths, n_cols).
 *   path_slot : NULL or DEVICE int32 [n, n_cols]: slot j of each column, -1 if the column has no path.
 */
int dmk_channels_td(const dmk_desc *desc,
                    const float *power_dbw, const float *phase_deg, const float *delay_s,
                    const float *aoa_az_deg, const float *aoa_el_deg,
                    const float *aod_az_deg, const float *aod_el_deg,
                    const double *ue_rot_deg, const float *doppler_hz,
                    int64_t n_users, int32_t ld,
                    void *out_c64,
                    uint8_t *fov_mask, uint8_t *valid_mask, int32_t *path_slot,
                    void *cuda_stream);

/* Time-domain channels plus the per-slot delays: the (a, tau) pair of the reference's Sionna adapter
 * (deepmimo/integrations/sionna_adapter.py:174-200).  `out_c64` [n, M_r, M_t, P] is, per user, exactly the block the adapter
 * copies into a[i_rx, :, i_tx, :, :, 0];
 *   tau : NULL or DEVICE float32 [n, P]: tau[u, j] = delay of the j-th valid path of user u, 0 in the empty slots
 *         (== tau[i_rx, i_tx, :num_paths] = ToA, :196-198).  Everything else as dmk_channels_td. */
int dmk_channels_td_tau(const dmk_desc *desc,
                        const float *power_dbw, const float *phase_deg, const float *delay_s,
                        const float *aoa_az_deg, const float *aoa_el_deg,
                        const float *aod_az_deg, const float *aod_el_deg,
                        const double *ue_rot_deg, const float *doppler_hz,
                        int64_t n_users, int32_t ld,
                        void *out_c64,
                        uint8_t *fov_mask, uint8_t *valid_mask, int32_t *path_slot, float *tau,
                        void *cuda_stream);

/* Beam amplitude map, fused (docs/manual.ipynb cell 105 of the reference):
 *   mean_abs[u, b] = mean over RX elements r and selected subcarriers k of | sum_t beams[b, t] H[u, r, t, k] |
 * == np.abs(F1 @ dataset.channel).mean(axis=1).mean(axis=-1) with F1 = beams (rows e.g. from steering_vec,
 * deepmimo/generator/geometry.py:322-339).  H is never written: the codebook is folded into the TX steering of
 * every path and the [n, M_r, n_beams, K] product is reduced in registers.  Frequency domain, no time axis, rx_filter = 0.
 *   beams_c64 : DEVICE complex64 [n_beams, M_t], row-major (M_t = bs_shape[0] * bs_shape[1], element order y fastest)
 *   mean_abs  : DEVICE float32 [n, n_beams]
 * masks as in dmk_channels_fd. */
int dmk_beam_amplitude_fd(const dmk_desc *desc,
                          const float *power_dbw, const float *phase_deg, const float *delay_s,
                          const float *aoa_az_deg, const float *aoa_el_deg,
                          const float *aod_az_deg, const float *aod_el_deg,
                          const double *ue_rot_deg,
                          int64_t n_users, int32_t ld,
                          const void *beams_c64, int32_t n_beams,
                          float *mean_abs,
                          uint8_t *fov_mask, uint8_t *valid_mask, uint8_t *clip_mask,
                          void *cuda_stream);

/* Per-path by-products of the prologue (the Dataset caches the reference fills lazily):
 *   angles_rot : NULL or DEVICE double [4, n, n_cols] = _aod_el_rot, _aod_az_rot, _aoa_el_rot, _aoa_az_rot
 *                (radians, before FoV NaN-ing; dataset.py:351-356)
 *   power_gain : NULL or DEVICE double [n, n_cols] = _power_linear_ant_gain (dataset.py:665-691)
 */
int dmk_path_prologue(const dmk_desc *desc,
                      const float *power_dbw,
                      const float *aoa_az_deg, const float *aoa_el_deg,
                      const float *aod_az_deg, const float *aod_el_deg,
                      const double *ue_rot_deg,
                      int64_t n_users, int32_t ld,
                      double *angles_rot, double *power_gain, uint8_t *fov_mask,
                      void *cuda_stream);

/* Per-user by-products of the same prologue (the lazy Dataset keys of the reference):
 *   num_paths            : NULL or DEVICE int32 [n]: paths left after FoV filtering, counted over all columns (dataset.py:613-619)
 *   los                  : NULL or DEVICE int32 [n]: 1 LoS / 0 NLoS / -1 no path, from `inter` of the first in-FoV path (dataset.py:569-611)
 *   pathloss_coherent    : NULL or DEVICE float32 [n]: -10 log10 |sum_p sqrt(p_lin) e^{j phase}|^2, NaN where the sum is 0 (dataset.py:541-566)
 *   pathloss_noncoherent : NULL or DEVICE float32 [n]: the same with the phases dropped (coherent=False)
 *   inter                : DEVICE float32 [n, n_cols] (ld) interaction codes (0 = line of sight); required when `los` is requested
 */
int dmk_user_byproducts(const dmk_desc *desc,
                        const float *power_dbw, const float *phase_deg,
                        const float *aoa_az_deg, const float *aoa_el_deg,
                        const float *aod_az_deg, const float *aod_el_deg,
                        const float *inter, const double *ue_rot_deg,
                        int64_t n_users, int32_t ld,
                        int32_t *num_paths, int32_t *los,
                        float *pathloss_coherent, float *pathloss_noncoherent,
                        void *cuda_stream);

/* Test hook for rounding point R2: the device restatement of NumPy's
 * float32 sin/cos used at geometry.py:301-302.  x, s, c are DEVICE float32 [n]. */
int dmk_np_sincosf(const float *x, float *s, float *c, int64_t n, void *cuda_stream);

const char *dmk_last_error(void);     /* thread-local text of the last error                  */
int         dmk_abi_version(void);    /* == DMK_ABI_VERSION                                   */
int64_t     dmk_launch_count(void);   /* kernels launched by this library since load (process-wide) */
const char *dmk_last_kernel(void);    /* name/variant of the last channel kernel launched     */

#ifdef __cplusplus
}
#endif
#endif /* DMK_H */
