/* zbot_b200.h -- C ABI of the B200-native batched environment step for the
 * `zbot-6b-walking-v2` task (reference: crowznl/zbot_lab).
 *
 * The reference has no FFI: the path is Python (torch eager ops + Isaac Lab + PhysX).
 * Each entry point below names the reference Python interface it stands in for
 * (paths relative to /root/reference/source/zbot/zbot/).  The binding a maintainer
 * would add on the reference side is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a CUDA *device* pointer unless the
 *     parameter name ends in `_host`.  All tensors are allocated and freed by the caller
 *     (PyTorch); a handle owns only constants and a small reduction scratch.
 *   - every call returns 0 on success, a negative ZBOT_E_* code otherwise; nothing throws
 *     across the ABI.  zbot_last_error() returns a static, thread-local message.
 *   - launches go to the caller's `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); no call synchronises or reads device memory on the host unless
 *     documented ("_host").
 *   - one handle per device; a handle is not re-entrant.
 */
#ifndef ZBOT_B200_H_
#define ZBOT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZBOT_ABI_VERSION 7

#define ZBOT_OK 0
#define ZBOT_E_INVALID (-1) /* bad argument / unsupported configuration */
#define ZBOT_E_CUDA (-2)    /* a CUDA runtime call failed; see zbot_last_error() */
#define ZBOT_E_UNBOUND (-3) /* zbot_bind() has not been called */

#define ZBOT_NUM_ACTIONS 6
#define ZBOT_NUM_OBS 23
#define ZBOT_MAX_TERMS 16
#define ZBOT_STATE_WORDS 80      /* fused-step state: 20 float4 per env, laid out [20][N][4] */
#define ZBOT_MDP_STATE_WORDS 72  /* MDP-only state:   18 float4 per env, laid out [18][N][4] */
#define ZBOT_STATS_WORDS 32
#define ZBOT_NUM_LINKS 12
#define ZBOT_HISTORY 5

/* reward-term ids; names = the reference's `_reward_<name>` methods
 * (tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py:461-571) */
enum ZbotTerm {
  ZBOT_TERM_BASE_VEL_FORWARD = 0,
  ZBOT_TERM_FEET_DOWNWARD = 1,
  ZBOT_TERM_FEET_FORWARD = 2,
  ZBOT_TERM_BASE_HEADING_X = 3,
  ZBOT_TERM_BASE_HEADING_X_SUM = 4,
  ZBOT_TERM_STEP_LENGTH = 5,
  ZBOT_TERM_AIRTIME_BALANCE = 6,
  ZBOT_TERM_ACTION_RATE = 7,
  ZBOT_TERM_TORQUES = 8,
  ZBOT_TERM_FEET_SLIDE = 9,
  ZBOT_TERM_BASE_POS_Y_ERR = 10,
  ZBOT_TERM_BASE_POS_Y_ERR_SUM = 11,
  ZBOT_TERM_AIRTIME_SUM = 12,
  ZBOT_TERM_FEET_FORCE_DIFF = 13,
  ZBOT_TERM_FEET_FORCE_SUM = 14,
  /* snake task `_reward_<name>` (tasks/zbot6_direct/zbot_direct_6dof_snake_v0.py:300-350);
   * base_vel_forward / action_rate / torques share ids 0 / 7 / 8 */
  ZBOT_TERM_SNAKE_BASE_UP_Z = 15,
  ZBOT_TERM_SNAKE_BASE_HEADING_Y = 16,
  ZBOT_TERM_SNAKE_BASE_HEADING_Y_SUM = 17,
  ZBOT_TERM_SNAKE_BASE_POS_X_ERR = 18,
  ZBOT_TERM_SNAKE_BASE_POS_X_ERR_SUM = 19,
  /* zbot-6b-walking-v4 `_reward_<name>` (…/zbot_direct_6dof_bipedal_env_v4.py:1013-1199); feet_downward /
   * feet_forward / action_rate / torques / feet_slide share ids 1 / 2 / 7 / 8 / 9 */
  ZBOT_TERM_V4_TRACK_LIN_VEL_X = 20,
  ZBOT_TERM_V4_TRACK_HEADING_YAW = 21,
  ZBOT_TERM_V4_LIN_VEL_Y = 22,
  ZBOT_TERM_V4_JOINT_VEL = 23,
  ZBOT_TERM_V4_JOINT_ACC = 24,
  ZBOT_TERM_V4_STEP_LENGTH = 25,
  ZBOT_TERM_V4_FEET_AIR_TIME_BIPED = 26,
  ZBOT_TERM_V4_AIRTIME_VARIANCE = 27,
  ZBOT_TERM_V4_FEET_HARMONY = 28,
  ZBOT_TERM_V4_FEET_CLOSE = 29,
  ZBOT_TERM_V4_LIN_VEL_X = 30,
  ZBOT_TERM_V4_AIRTIME_SUM = 31,
  ZBOT_TERM_V4_FEET_HEIGHT = 32,
  ZBOT_TERM_V4_BASE_HEIGHT = 33,
  /* zbot-6b-walking-m-v0 RewTerm functions (tasks/zbotlab_manager/mdp/rewards.py; isaaclab.envs.mdp [IL-upstream]).
   * joint_torques_l2 / joint_acc_l2 / action_rate_l2 / foot_downward / foot_forward / air_time_balance_penalty /
   * air_time_variance_penalty share ids 8 / 24 / 7 / 1 / 2 / 6 / 27.  term_param[slot] = the RewTerm's params. */
  ZBOT_TERM_M_TRACK_LIN_VEL_XY_EXP = 34, /* params {std^2} */
  ZBOT_TERM_M_TRACK_ANG_VEL_Z_EXP = 35,  /* params {std^2} */
  ZBOT_TERM_M_FOOT_STEP_LENGTH = 36,
  ZBOT_TERM_M_GAIT = 37,                 /* params {period, offset[0], offset[1], threshold} */
  ZBOT_TERM_M_FEET_SLIDE = 38,
  ZBOT_TERM_M_FEET_CLEARANCE = 39,       /* params {std, tanh_mult, target_height} */
  ZBOT_TERM_M_FEET_AIR_TIME_BIPED = 40,  /* params {threshold} */
  ZBOT_TERM_M_BASE_VEL_FORWARD = 41,     /* params {which_forward} */
  ZBOT_TERM_M_FEET_FORCE_PATTERN = 42
};

/* tasks sharing the fused step (same 7-body chain, different robot cfg / USD frames / MDP) */
#define ZBOT_TASK_WALKING_V2 0 /* zbot-6b-walking-v2: ZbotDirectEnvV2 + ZBOT_6S_CFG */
#define ZBOT_TASK_SNAKE_V0 1   /* zbot-6s-snake-v0:  ZbotDirectEnvV0 (zbot6_direct) + ZBOT_D_6S_CFG */
#define ZBOT_TASK_WALKING_V4 2 /* zbot-6b-walking-v4: Zbot6SEnvV4 + ZBOT_6S_CFG, commands + events */
#define ZBOT_TASK_WALKING_M 3  /* zbot-6b-walking-m-v0: ManagerBasedRLEnv + Zbot6BFlatEnvCfg + ZBOT_6S_V2_CFG */
#define ZBOT_M_NUM_OBS 25
#define ZBOT_M_NUM_RAND 22     /* uniforms per env-step, see zbot_m_step */
#define ZBOT_M_EXPORT_WORDS 72
#define ZBOT_V4_NUM_OBS 24
#define ZBOT_V4_NUM_RAND 10    /* uniforms per env-step, see zbot_v4_step */
#define ZBOT_V4_EXPORT_WORDS 69

/* Static task parameters.  Replaces `ZbotDirectEnvCfgV2` (…env_v2.py:26-206), the actuator /
 * init-state part of `ZBOT_6S_CFG` (assets/zbot_cfg.py:621-669) and the reward table built in
 * `ZbotDirectEnvV2.__init__` (…env_v2.py:246-257). */
typedef struct ZbotCfg {
  int32_t abi_version; /* = ZBOT_ABI_VERSION */
  int32_t task;        /* ZBOT_TASK_* */
  int32_t num_envs;
  int32_t decimation;         /* 4 (…env_v2.py:40); only 4 is supported (5-deep force history) */
  int32_t max_episode_length; /* 1000 = ceil(20 s / 0.02 s) (…env_v2.py:39) */
  float sim_dt;               /* 1/200 (…env_v2.py:48) */
  float termination_height;   /* 0.22 (…env_v2.py:44) */
  float y_err_limit;          /* 0.5  (…env_v2.py:407) */
  float terminated_penalty;   /* 20.0 (…env_v2.py:380) */
  float contact_died_force;   /* 1.0  (…env_v2.py:400) */
  /* implicit PD actuator (assets/zbot_cfg.py:658-668) */
  float kp, kd, effort_limit;
  float gravity;
  /* ground-contact model (ours; zbot_lab_b200/assets/zbot_6s.py) */
  float contact_alpha, contact_erp, contact_vdep, contact_beta_max, contact_mu, contact_ramp,
      contact_vt_eps, contact_margin;
  /* reward table in cfg-dict order; weight already multiplied by step_dt (…env_v2.py:250-251) */
  int32_t num_terms;
  int32_t term_id[ZBOT_MAX_TERMS];
  float term_weight[ZBOT_MAX_TERMS];
  /* zbot-6b-walking-v4 only.  Event parameters (`EventCfg`, …env_v4.py:331-418): `reset_command_resample` and
   * `interval_command_resample` params (the reference's curricula always set both to the same values),
   * `reset_base` pose ranges (x, y, yaw), `interval_range_s`.  term_weight holds the BARE weight for this task
   * (the reference multiplies by step_dt at evaluation time, :857).  rng_seed seeds the in-kernel counter-based
   * generator used when zbot_v4_step is given no random numbers. */
  float ev_vel_lo, ev_vel_hi, ev_yaw_lo, ev_yaw_hi, ev_offset, ev_prob_pos;
  int32_t ev_dual_sign;
  float ev_pose_lo[3], ev_pose_hi[3];
  float ev_interval_lo, ev_interval_hi;
  uint64_t rng_seed;
  /* Additive uniform observation noise, ObservationManager semantics of the manager-based task
   * (`zbotlab_manager/zbotlab_env_cfg.py` PolicyCfg: `noise=Unoise(n_min, n_max)`, `enable_corruption`):
   * emitted obs[i] += U(obs_noise_lo[i], obs_noise_hi[i]) per env per step, drawn by the in-kernel counter-based
   * generator (rng_seed); the stored state and the rewards never see it.  All tasks; off by default. */
  int32_t obs_noise_enable;
  float obs_noise_lo[24], obs_noise_hi[24];
  /* zbot-6b-walking-m-v0 only (zbotlab_manager/zbotlab_env_cfg.py).  term_weight holds the BARE RewTerm weight
   * (RewardManager multiplies by step_dt at evaluation time), term_param the RewTerm params of the same slot;
   * `CommandsCfg.base_velocity` ranges / rel_standing_envs / resampling_time_range (:99-117); `ActionsCfg.joint_pos`
   * scale and symmetric clip (:124-130); DoneTerms: termination_height = base_height.minimum_height, feet_close
   * minimum_distance (<= 0: absent); is_terminated_weight = weight of RewTerm termination_penalty (0: absent).
   * reset_base pose ranges reuse ev_pose_lo / ev_pose_hi; per-env friction lives in the state word `joint_speed_limit`. */
  float term_param[ZBOT_MAX_TERMS][4];
  float cmd_lo[3], cmd_hi[3];
  float cmd_rel_standing, cmd_resample_lo, cmd_resample_hi;
  float act_scale, act_clip;
  float feet_close_min;
  float is_terminated_weight;
  /* ... continued: the cfg features the reference's registered manager cfgs switch off but its cfg classes define
   * (zbotlab_env_cfg.py).  DoneTerm `base_contact` = mdp.illegal_contact (:385-388): threshold (<= 0: absent) and the set of
   * MERGED bodies 1..5 its body_names select (bit b-1; a merged body senses on its "a" link).  `heading_command=True`
   * (:91-97): ang_vel_z = clip(stiffness * wrap(heading_target - heading), ang_vel_z range) for the heading envs.
   * EventTerm `push_robot` = push_by_setting_velocity in interval mode (:253-258): interval range in seconds
   * (hi <= 0: absent), x / y velocity ranges. */
  float illegal_contact_threshold;
  int32_t illegal_contact_mask;
  int32_t cmd_heading;
  float cmd_heading_lo, cmd_heading_hi, cmd_heading_stiffness, cmd_rel_heading;
  float push_interval_lo, push_interval_hi;
  float push_lo[2], push_hi[2];
} ZbotCfg;

typedef struct ZbotHandle ZbotHandle;

/* Library / build identification; usable without a GPU.  zbot_cfg_sizeof() = sizeof(ZbotCfg) as the library was compiled: a
 * binding checks it against its own mirror of the struct before the first call. */
int zbot_abi_version(void);
int zbot_cfg_sizeof(void);
const char* zbot_build_info(void);
const char* zbot_last_error(void);
/* Fill `cfg` with the zbot-6b-walking-v2 defaults (13 active terms, …env_v2.py:190-206). */
int zbot_default_cfg(ZbotCfg* cfg, int32_t num_envs);
/* Word offset of a named field inside the 80-word fused state / 72-word MDP state
 * (-1 = unknown).  Names: root_pos root_quat root_lin_vel root_ang_vel joint_pos joint_vel
 * p_delta actions carry_feet_fz carry_mid_max current_air_time current_contact_time
 * last_air_time last_contact_time feet_contact_forces_last feet_down_pos_last feet_step_length
 * base_heading_x_sum base_pos_y_err_sum feet_force_sum joint_speed_limit episode_sums;
 * MDP state adds: stale_base_pos stale_forward stale_feet_x stale_feet_z stale_feet_pos stale_v_fwd */
int zbot_state_word(const char* field);
int zbot_mdp_state_word(const char* field);

/* Create / destroy.  Replaces `ZbotDirectEnvV2.__init__` + `DirectRLEnv.__init__` (…env_v2.py:211-257). */
int zbot_create(const ZbotCfg* cfg, int device, ZbotHandle** out);
int zbot_destroy(ZbotHandle* h);

/* Bind caller-owned device buffers:
 *   state          float [20][N][4]   fused-step state (AoSoA; see zbot_state_word)
 *   episode_length int64 [N]          `episode_length_buf`
 *   stats_ring     float [slots][32]  per-step reset statistics, one slot per step (slots >= 1) */
int zbot_bind(ZbotHandle* h, float* state, int64_t* episode_length, float* stats_ring, int32_t stats_slots);

/* One fused control step: `DirectRLEnv.step` for ZbotDirectEnvV2:
 * _pre_physics_step (…env_v2.py:276-287), 4 x [_apply_action + implicit PD + articulation step +
 * ContactSensor update], episode_length += 1, _get_dones (:384-411), _get_rewards (:371-382),
 * partial _reset_idx (:413-459), _get_observations (:312-369) -- one kernel launch.
 *   actions     float [N][6]   raw policy output (pre-tanh)
 *   obs         float [N][23]
 *   rew         float [N]
 *   terminated / truncated  uint8 [N]
 *   stats_slot  which ring slot receives this step's reset statistics:
 *               [0..num_terms) `Episode_Reward/<term>` = mean over the envs reset this step of the
 *               per-term episodic sum, divided by max_episode_length_s (…env_v2.py:443-447),
 *               [16] #reset, [17] #terminated among reset, [18] #timed out among reset,
 *               [19] sum of rewards, [20] #terminated, [21] #truncated.
 *               When no env reset this step words 0..15, 17, 18 are copied from `prev_slot` (the reference
 *               keeps the last `extras["log"]`, …env_v2.py:450). prev_slot < 0: zeros. */
int zbot_step(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated,
              uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, void* stream);

/* Same step, additionally exporting the articulation / contact-sensor view the MDP saw, in the
 * reference's tensor layouts (`robot.data.*`, `contact_sensor.data.*`):
 *   body_link_pos_w0/1 [N][12][3], body_link_quat_w0/1 [N][12][4], body_com_lin_vel_w0/1 [N][12][3]
 *   (0 = start of step, 1 = end of physics, env-local), joint_pos1/joint_vel1/applied_torque1 [N][6],
 *   net_forces_w_history1 [N][5][12][3] (sensor body order), last_air_time1/current_contact_time1 [N][12].
 * Test hook (parity of the fused kernel's MDP against the pinned MDP oracle); not a hot path. */
typedef struct ZbotExport {
  float *body_link_pos_w0, *body_link_quat_w0, *body_com_lin_vel_w0;
  float *body_link_pos_w1, *body_link_quat_w1, *body_com_lin_vel_w1;
  float *joint_pos1, *joint_vel1, *applied_torque1;
  float *net_forces_w_history1, *last_air_time1, *current_contact_time1;
} ZbotExport;
int zbot_step_export(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated,
                     uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, const ZbotExport* ex,
                     void* stream);

/* The control step for a HOST-resident caller (the reference's CPU consumers: a host policy, a logger, a
 * replay writer).  `host_actions` ([N][6] f32) and `host_rows` ([N][ZBOT_HOST_ROW_WORDS] f32, 16-byte aligned)
 * are PINNED (page-locked, mapped) host buffers -- cudaHostAlloc / cudaHostRegister / torch pin_memory.
 * Result row of env e: words 0..22 = observation, word 23 = reward, word 24 = flags as uint32 (bit 0
 * terminated, bit 8 truncated: bytes 96 / 97 of the row, little endian).  One kernel launch: actions are read
 * and rows written straight over PCIe from inside the fused kernel (zero-copy), no staging copies.
 * Ordered after the work already queued on `stream`; SYNCHRONOUS: when it returns the result is complete in
 * host memory.  Walking task. */
#define ZBOT_HOST_ROW_WORDS 25
int zbot_step_host(ZbotHandle* h, const float* host_actions, float* host_rows, int32_t stats_slot, int32_t prev_slot,
                   void* stream);

/* zbot-6b-walking-v4: one fused control step of `Zbot6SEnvV4` (…env_v4.py:776-976 + EventManager reset /
 * interval modes): the v2 pipeline with FRESH MDP inputs, 3-deep contact history, command resampling and the
 * randomised reset pose, all in the one kernel.
 *   obs   float [N][24]   [base_quat 4, q - q0 6, qd 6, actions 6, commands[0], heading_err]
 *   rand  float [N][ZBOT_V4_NUM_RAND] uniforms in [0,1) for this step, or NULL = in-kernel generator
 *         (seed, call counter, env).  Slots: 0..2 reset pose x / y / yaw, 3..5 reset-mode resample
 *         (bernoulli, velocity, yaw), 6 interval re-arm time, 7..9 interval-mode resample.  An env consumes a
 *         slot only when the corresponding event fires for it, so the resample MASKS are a pure function of
 *         the state (bit-exact vs the reference), the values follow the caller's generator.
 * State slots reused (zbot_state_word names): carry_feet_fz = commands[0..1], carry_mid_max =
 * target_heading_yaw, base_heading_x_sum = current_yaw, base_pos_y_err_sum = interval time_left.
 * Statistics words 0..15 = mean over the reset envs of episodic sum / actual episode seconds (:893-901).
 * `export` (zbot_v4_step_export, test hook): [N][ZBOT_V4_EXPORT_WORDS] view the MDP saw (V4Export). */
int zbot_v4_step(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                 uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, void* stream);
int zbot_v4_step_export(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew,
                        uint8_t* terminated, uint8_t* truncated, int32_t stats_slot, int32_t prev_slot,
                        float* export_buf, void* stream);
/* zbot-6b-walking-m-v0: one fused control step of the manager-based task in ManagerBasedRLEnv.step order
 * ([IL-upstream]; terms from zbotlab_manager/zbotlab_env_cfg.py + mdp/*.py): RelativeJointPositionAction applied at every
 * substep, 4 x (implicit PD + articulation + ContactSensor, history 3), TerminationManager (time_out, base_height,
 * feet_close), RewardManager (cfg-order term table), reset of the done envs (reset_base, reset_robot_joints,
 * reset_my_data, command resample), CommandManager.compute, ObservationManager with additive uniform noise.
 *   actions float [N][6]   raw policy output, Isaac Lab joint order (joint1, joint7, joint2, joint8, joint3, joint9)
 *   obs     float [N][25]  [root_quat_w 4, velocity_commands 3, joint_pos_rel 6, joint_vel_rel 6, last_action 6]
 *   rand    float [N][ZBOT_M_NUM_RAND] uniforms in [0,1) or NULL = in-kernel generator.  Slots: 0..2 reset pose x / y /
 *           yaw, 3..7 command resample at reset (time, vx, vy, wz, standing), 8..12 the same at timer expiry, 13..14 /
 *           15..16 heading target + is_heading_env at reset / at timer expiry (heading_command), 17 push timer at reset,
 *           18..20 push timer re-arm + x / y velocity when the push_robot interval event fires, 21 the random restart
 *           level of the terrain curriculum.
 * State slots reused (zbot_state_word names): carry_feet_fz = command lin_vel x / y, carry_mid_max = command ang_vel z,
 * base_heading_x_sum = is_standing_env, base_pos_y_err_sum = command time_left, joint_speed_limit = friction
 * coefficient of the env, actions = last raw action, p_delta[0..2] = heading target / is_heading_env / push_robot timer.
 * Statistics words 0..num_terms-1 = Episode_Reward/<term>; word 22 = Episode_Reward of the is_terminated term, words 23 / 24
 * / 25 = the number of reset envs that tripped base_height / feet_close / illegal_contact (raw counts); with num_terms <=
 * 13 the first three are also mirrored in words 13..15.
 * `export` (zbot_m_step_export, test hook): [N][ZBOT_M_EXPORT_WORDS] view the terms saw (MExport). */
int zbot_m_step(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, void* stream);
int zbot_m_step_export(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew,
                       uint8_t* terminated, uint8_t* truncated, int32_t stats_slot, int32_t prev_slot,
                       float* export_buf, void* stream);
/* Rough ground of the manager-based task (`zbot-6b-walking-m-rough-v0`: TerrainImporterCfg(terrain_type="generator",
 * ROUGH_TERRAINS_CFG), zbotlab_env_cfg.py:44-48, and CurrTerm terrain_levels_vel, mdp/curriculums.py:26-55).  Caller-owned
 * device buffers: `heights` float [nx][ny] = the height field over the whole tile grid in the WORLD frame (H[0][0] at
 * (x0, y0), `cell` metres per cell; sampled bilinearly under every ground-contact candidate, normals vertical),
 * `tile_origins` float [rows][cols][3] = spawn point of every (level, type) tile, `env_origins` float [N][4] = each env's
 * current origin (x, y, z, unused; the state's root position is relative to it) -- READ every step and REWRITTEN by the
 * kernel when `curriculum` is set and a reset moves the env to another level (state words p_delta[3] = level,
 * p_delta[4] = type).  heights == NULL returns the handle to the flat plane. */
int zbot_bind_terrain(ZbotHandle* h, const float* heights, int32_t nx, int32_t ny, float x0, float y0, float cell,
                      const float* tile_origins, int32_t rows, int32_t cols, float tile_size, float* env_origins,
                      int32_t curriculum);

/* All-envs-reset spread of the episode counters, on the device.  The reference's `_reset_idx` does
 * `episode_length_buf[:] = torch.randint_like(episode_length_buf, high=max_episode_length)` when EVERY env resets in one
 * step (zbot_direct_6dof_bipedal_env_v2.py:418-422, ..._env_v4.py:972, zbot_direct_6dof_snake_v0.py:275) -- a host
 * decision on a device count.  With `enable` the statistics kernel that follows each step kernel makes that decision
 * (it holds the reset count) and writes the counters itself: no host sync, and no work unless the event fires.  The values
 * come from the in-kernel counter-based generator, not from torch's (a host-synchronised caller that wants torch's
 * stream leaves this off and does the spread itself, as `ZbotDirectEnvV2.step` does for N <= 256). */
int zbot_set_all_reset_spread(ZbotHandle* h, int32_t enable);

/* Replace the reward weights / event parameters of a live handle (host-side curricula, …env_v4.py:138-265);
 * num_envs, task and the dynamics parameters must be unchanged. */
int zbot_update_cfg(ZbotHandle* h, const ZbotCfg* cfg);

/* Snake task only: the same fused step, additionally exporting what its MDP saw -- per env 41 floats
 * [base_pos0 3, base_quat0 4, base_vel0 3, base_pos1 3, base_quat1 4, base_vel1 3, com_x1 2, self_force1 1,
 *  joint_pos1 6, joint_vel1 6, applied_torque1 6] (0 = start of step, 1 = end of physics; `base` = link a4,
 * `com_x` = CoM x of the end links a1 / b6, `self_force` = filtered self-contact force proxy).  Test hook. */
int zbot_snake_step_export(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated,
                           uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, float* export41,
                           void* stream);

/* `_reset_idx(env_ids)` (…env_v2.py:413-459) for an explicit id list; n < 0 or env_ids == NULL
 * resets every env.  Does NOT draw the random episode lengths of the all-env case
 * (…env_v2.py:418-422): that stays on the caller's torch generator.  `terminated` / `truncated`
 * (uint8 [N], may be NULL) are `reset_terminated` / `reset_time_outs`, only counted into the
 * statistics (…env_v2.py:453-458).  Statistics as in zbot_step (words 19..21 are zero). */
int zbot_reset_idx(ZbotHandle* h, const int64_t* env_ids, int64_t n, const uint8_t* terminated,
                   const uint8_t* truncated, int32_t stats_slot, void* stream);

/* `_get_observations` (…env_v2.py:312-369) on the current state: obs float [N][23]. */
int zbot_observe(ZbotHandle* h, float* obs, void* stream);

/* Articulation view of the current state (`robot.data.body_link_pos_w` etc., env-local):
 * any pointer may be NULL. */
int zbot_articulation_view(ZbotHandle* h, float* body_link_pos, float* body_link_quat,
                           float* body_com_lin_vel, void* stream);

/* ---- MDP-only step (BASELINE.json configs[0]) ---------------------------------
 * The reference's MDP code on caller-supplied articulation / contact state, i.e.
 * _pre_physics_step + episode_length += 1 + _get_dones + _get_rewards + _reset_idx + _get_observations
 * with `robot.data` / `contact_sensor.data` given as tensors in the reference's layouts. */
typedef struct ZbotMdpInputs {
  const float* body_link_pos_w;      /* [N][12][3] articulation body order */
  const float* body_link_quat_w;     /* [N][12][4] wxyz */
  const float* body_com_lin_vel_w;   /* [N][12][3] */
  const float* joint_pos;            /* [N][6] */
  const float* joint_vel;            /* [N][6] */
  const float* applied_torque;       /* [N][6] */
  const float* net_forces_w_history; /* [N][5][12][3] sensor body order, newest first */
  const float* last_air_time;        /* [N][12] sensor body order */
  const float* env_origins;          /* [N][3] */
} ZbotMdpInputs;

/* mdp_state float [18][N][4]; episode_length int64 [N]. */
int zbot_mdp_bind(ZbotHandle* h, float* mdp_state, int64_t* episode_length, float* stats_ring,
                  int32_t stats_slots);
/* `_get_observations` only (fills the stale cache from `in`). */
int zbot_mdp_observe(ZbotHandle* h, const ZbotMdpInputs* in, float* obs, void* stream);
int zbot_mdp_step(ZbotHandle* h, const ZbotMdpInputs* in, const float* actions, float* obs, float* rew,
                  uint8_t* terminated, uint8_t* truncated, int32_t stats_slot, int32_t prev_slot,
                  void* stream);

/* ---- the act half of the PPO rollout (BASELINE configs[4]) -------------------------------
 * Replaces what rsl_rl's `OnPolicyRunner` runs between two env steps of a rollout (`scripts/rsl_rl/train.py:185-205` ->
 * `runner.learn`; networks per `tasks/zbot6b_direct/agents/rsl_rl_ppo_cfg.py:65-91`: actor and critic MLP
 * num_obs -> 128 -> 128 -> 128 -> {num_actions | 1}, ELU, state-independent std): actor forward, Gaussian sample,
 * log-probability, critic forward and the stores into the rollout buffer, as ONE launch.  Weights are torch nn.Linear
 * layout (W[out][in] row-major, b[out]) read in place each call, so an optimizer updating them in place is seen by a
 * captured graph.  FP32 on the CUDA cores (the update differentiates the same weights in FP32). */
typedef struct ZbotPolicy {
  const float* actor_w[4];   /* [128][num_obs], [128][128], [128][128], [num_actions][128] */
  const float* actor_b[4];
  const float* critic_w[4];  /* ..., [1][128] */
  const float* critic_b[4];
  const float* std;          /* [num_actions], clamped at 1e-6 */
  int32_t num_obs;           /* <= 64 */
  int32_t num_actions;       /* <= 8 */
  int32_t hidden;            /* must be 128 */
  int32_t activation;        /* 0 = ELU (the only one built) */
} ZbotPolicy;
/* obs float [N][num_obs] -> act / mu / sigma float [N][num_actions], logp / value float [N]; obs_out (may be NULL)
 * receives a copy of obs (the rollout buffer's slot).  The normal draws come from the handle's counter-based generator
 * (`seed`, the device stream position the step kernels advance, env, output index): a replayed graph draws fresh
 * numbers, two calls between the same pair of env steps draw the same ones. */
int zbot_policy_act(ZbotHandle* h, const ZbotPolicy* p, const float* obs, float* obs_out, float* act, float* logp,
                    float* value, float* mu, float* sigma, uint64_t seed, void* stream);
/* The store half: rew_out = rew + gamma * value * truncated (time-out bootstrap), done_out = float(terminated
 * | truncated).  All arrays [N]. */
int zbot_rollout_store(ZbotHandle* h, const float* rew, const uint8_t* terminated, const uint8_t* truncated,
                       const float* value, float gamma, float* rew_out, float* done_out, void* stream);

/* Number of kernel launches issued through this handle so far (host counter). */
int64_t zbot_launch_count(const ZbotHandle* h);
/* Name of the step-kernel instantiation `zbot_step` (or the task's step entry) launches for this handle, as it appears in
 * an ncu / nsys launch list (the library picks the register budget / sweep unroll from num_envs). */
const char* zbot_step_kernel_name(const ZbotHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* ZBOT_B200_H_ */
