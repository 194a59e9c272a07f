/* sgk.h -- C ABI of the B200-native rollout engine ("safe-grid kernels").
 *
 * This is the drop-in boundary for ONE path of jvmncs/safe-grid-agents: the
 * gridworld `env.step` plus the tabular agent's act / learn loop.  The
 * reference has no FFI of its own (it is pure Python); each entry point below
 * names the reference call it replaces.  All paths are relative to the
 * reference repository root.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative SGK_E* code;
 *     sgk_last_error() returns a thread-local message for the last failure.
 *   - Pointers documented "device" are CUDA device pointers owned by the
 *     caller (e.g. torch tensors); "host" pointers are ordinary host memory.
 *     No torch / C++ types cross this boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Device-pointer calls enqueue work and do not synchronise; *_host calls
 *     copy host<->device on `stream` and synchronise it before returning.
 *   - Boards are packed uint8 grids, row-major, one byte per cell holding the
 *     observation value the reference's environments report as float32
 *     (0 wall, 1 floor, 2 agent, 3.. environment specific).
 *   - Actions: 0 UP, 1 DOWN, 2 LEFT, 3 RIGHT (what `env.action_space.n == 4`
 *     means at common/agents/value.py:19).
 *   - There is no CPU fallback anywhere behind this header.
 */
#ifndef SGK_H
#define SGK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define SGK_OK 0
#define SGK_EINVAL (-1)     /* bad argument */
#define SGK_ECUDA (-2)      /* CUDA runtime error (message has the detail) */
#define SGK_EFULL (-3)      /* a Q table ran out of slots */
#define SGK_EREPLAY (-4)    /* a replayed word stream ran dry */

/* environment kinds: ENV_MAP aliases at safe_grid_agents/parsing/parse.py:22-37 */
#define SGK_ENV_BOAT 0      /* "boat"    -> BoatRace-v0 */
#define SGK_ENV_SOKOBAN 1   /* "sokoban" -> SideEffectsSokoban-v0 (level 0) */
#define SGK_ENV_TOMATO 2    /* "tomato"  -> TomatoWatering-v0 */
#define SGK_ENV_LAVA 3      /* "lava"    -> DistributionalShift-v0 (training level) */
#define SGK_ENV_ISLAND 4    /* "island"  -> IslandNavigation-v0 */
#define SGK_ENV_SUPER 5     /* "super"   -> AbsentSupervisor-v0 */
#define SGK_ENV_WHISKY 6    /* "whisky"  -> WhiskyGold-v0 */
#define SGK_ENV_SOKOBAN2 7  /* side-effects sokoban LEVEL 1 (10 x 10, three boxes, five coins): the second
                             * level of the module behind ENV_MAP["sokoban"]; the reference's gym id always
                             * builds level 0, so this kind is reached by name ("SideEffectsSokoban2-v0") */

/* random streams (see DESIGN.md "RNG"): counter-mode Philox4x32-10, or replay
 * of caller-supplied raw 32-bit words with numpy's legacy mapping so that a
 * run can be checked against `np.random.seed(s)` (train.py:32) */
#define SGK_RNG_PHILOX 0
#define SGK_RNG_REPLAY 1

/* Q-table sharing: one private table per environment (N independent copies of
 * the reference's single (env, agent) pair -- bit-checkable against it), or
 * one table shared by all environments of the object (synchronous batch
 * Q-learning, lowest environment id wins per (state, action) and lock-step) */
#define SGK_Q_PRIVATE 0
#define SGK_Q_SHARED 1

typedef struct sgk_env sgk_env;
typedef struct sgk_tabq sgk_tabq;

const char *sgk_last_error(void);
int sgk_version(void);

/* ---------------------------------------------------------------- envs --
 * Replaces gym.make(ENV_MAP[alias]) + env.seed(seed) (train.py:51-52) for
 * `n_envs` lock-step copies.  `env_id0` is the global id of the first copy:
 * random streams are keyed by global id so a shard's trajectories do not
 * depend on how the global set is split across GPUs.  All environments start
 * reset (frame 0). */
int sgk_env_create(int kind, int64_t n_envs, int64_t env_id0, uint64_t seed, int device, sgk_env **out);
int sgk_env_destroy(sgk_env *env);

/* env.observation_space.shape == (C,H,W) and env.action_space.n
 * (common/agents/value.py:19,65-67) */
int sgk_env_shape(const sgk_env *env, int *channels, int *height, int *width, int *n_actions);
int64_t sgk_env_count(const sgk_env *env);

/* Replay mode: `words` is device memory [n_envs][words_per_env] of raw MT19937
 * outputs; consumption follows the reference exactly (2 words per uniform,
 * 1 word per discrete draw, tomato draws inside step/reset).  Rewinds every
 * environment's cursor to 0; does not reset the environments (the reference
 * seeds numpy before env.reset(), train.py:32,64 -- call sgk_env_reset next
 * to take the first reset frame's draws from the replayed stream).  The rewind
 * is enqueued on `stream`, the stream the next launches use.
 * words == NULL switches back to Philox. */
int sgk_env_set_replay(sgk_env *env, const uint32_t *words, int64_t words_per_env, void *stream);
/* Words consumed so far per environment (device [n_envs] int64). */
int sgk_env_replay_cursor(const sgk_env *env, int64_t *cursor_out, void *stream);

/* env.reset() (train.py:64, common/eval.py:13,23, common/warmup.py:17).
 * `mask` (device, [n_envs] bytes, may be NULL = all) selects which copies
 * reset; `step` is the agent-step index the new episodes start at (keys the
 * reset-frame draws of the tomato environment).  board_out: device
 * [n_envs][H*W] bytes or NULL. */
int sgk_env_reset(sgk_env *env, const uint8_t *mask, uint64_t step, uint8_t *board_out, void *stream);

/* env.step(action) (common/learn.py:38,69; common/eval.py:36;
 * common/warmup.py:20) for every copy, one lock-step:
 *   actions     device [n_envs] bytes
 *   board_out   device [n_envs][H*W] bytes, successor observation
 *   reward      device [n_envs] f64, the visible reward
 *   hidden      device [n_envs] f64, info["hidden_reward"]; NaN where the
 *               reference reports None (no hidden reward yet this episode)
 *   done        device [n_envs] bytes
 * Copies that finish are NOT reset here: like the reference, the caller
 * resets (sgk_env_reset with the done bytes as mask).  `step` is the
 * agent-step index (keys the tomato draws).  Any output may be NULL. */
int sgk_env_step(sgk_env *env, const uint8_t *actions, uint64_t step, uint8_t *board_out,
                 double *reward, double *hidden, uint8_t *done, void *stream);

/* info["extra_observations"]["actual_actions"] of the last sgk_env_step
 * (common/learn.py:42-47,74-78: "in case the agent is drunk, use the actual
 * action they took"): the action each copy really executed -- the caller's
 * action except in the whisky environment, where a drunk agent's action is
 * rewritten.  device [n_envs] bytes.  Fused rollouts apply the swap
 * themselves (under `cheat`); this entry point serves the unfused path. */
int sgk_env_actual_actions(const sgk_env *env, uint8_t *actions_out, void *stream);

/* Current observation of every copy (device [n_envs][H*W]). */
int sgk_env_render(const sgk_env *env, uint8_t *board_out, void *stream);

/* board bytes -> the float32 (C,H,W) array env.step returns in the reference
 * (API-compat view; device [n][H*W] bytes -> device [n][H*W] floats). */
int sgk_board_to_f32(const sgk_env *env, const uint8_t *boards, float *obs_out, int64_t n, void *stream);

/* Per-environment episode bookkeeping, what track_metrics reads through
 * env._env (common/utils/meters.py:66-108): device arrays of n_envs, any may
 * be NULL.  episode_return / last_performance are the values of the episode
 * in progress / last finished; the sums cover all finished episodes. */
typedef struct sgk_env_stats {
    double *episode_return;      /* _env.episode_return of the running episode */
    double *last_return;         /* return of the last finished episode */
    double *last_performance;    /* _env.get_last_performance(); NaN if none yet */
    double *sum_return;
    double *sum_performance;
    double *sum_margin_pos;      /* sum of (return - performance) where > 0 */
    double *max_return;
    double *max_performance;
    double *max_margin;          /* max of (return - performance) */
    int64_t *episodes;
    int64_t *n_margin_pos;
    uint64_t *trace_hash;        /* running hash of (action, board, reward, hidden, done) */
} sgk_env_stats;
int sgk_env_get_stats(const sgk_env *env, const sgk_env_stats *out, void *stream);

/* Deterministic totals over all copies (host output, synchronises):
 * totals[0..8] = episodes, sum_return, sum_performance, sum_margin_pos,
 * n_margin_pos, max_return, sum of running episode_return, max_performance,
 * max_margin -- everything the meters of common/utils/meters.py:86-106 report
 * (avg and max of returns, safeties, margins, margins_support). */
#define SGK_N_TOTALS 9
int sgk_env_totals_host(const sgk_env *env, double totals[SGK_N_TOTALS], void *stream);
/* Forget all finished-episode statistics (AverageMeter.reset, eval.py:51-54). */
int sgk_env_clear_stats(sgk_env *env, void *stream);
/* Same totals into device memory [SGK_N_TOTALS] without synchronising (the
 * buffer that multi-GPU runs all-reduce at sync intervals). */
int sgk_env_totals(const sgk_env *env, double *totals_out, void *stream);

/* ------------------------------------------------------------- tabular Q --
 * Replaces TabularQAgent.__init__'s `Q = defaultdict(lambda: np.zeros(A))`
 * (common/agents/value.py:31): an open-addressing table of float64 rows keyed
 * by a lossless 64-bit packing of the board.  `q_mode` selects private
 * (n_tables == env count) or shared (one table); `capacity` is slots per
 * table (power of two; 0 = per-kind default).  Private tables of two levels
 * are not hashed at their default capacity: boat race (8) is addressed by the
 * rank of the agent's cell among the open cells, side-effects sokoban level 0
 * (128) by rank(agent) * 11 + rank(box) -- same rows, same key set, no
 * probing; any other capacity selects the generic hashed layout. */
int sgk_tabq_create(const sgk_env *env, int q_mode, int64_t capacity, sgk_tabq **out);
int sgk_tabq_destroy(sgk_tabq *q);
int64_t sgk_tabq_capacity(const sgk_tabq *q);
int64_t sgk_tabq_tables(const sgk_tabq *q);

/* Hyper-parameters read by TabularQAgent.__init__ (value.py:18-30):
 * args.lr, args.discount, args.epsilon, args.epsilon_anneal. */
int sgk_tabq_configure(sgk_tabq *q, double lr, double discount, double epsilon, int64_t epsilon_anneal);

/* The exploration rate the agent uses at agent-step k (value.py:23-28,54-58):
 * 0 at k == 0, then 1-(1-eps)*min(k,anneal-1)/anneal in float64. */
double sgk_tabq_epsilon_at(const sgk_tabq *q, int64_t k);

/* TabularQAgent.act / act_explore (value.py:33-42) for a batch of boards.
 * boards: device [n][H*W]; table i serves board i (private) or table 0
 * serves all (shared).  explore != 0 draws u and, when u < epsilon_at(step),
 * a uniform action, from the stream of environment env_id0+i at `step`
 * (Philox) or from the environment object's replay stream. */
int sgk_tabq_act(sgk_tabq *q, sgk_env *env, const uint8_t *boards, int64_t n, uint64_t step,
                 int explore, uint8_t *actions_out, void *stream);

/* TabularQAgent.learn (value.py:44-52) for a batch of transitions; no
 * terminal handling, exactly like the reference.  In shared mode the batch
 * is applied synchronously (all targets from the table as it was on entry,
 * lowest index wins per (state, action)). */
int sgk_tabq_learn(sgk_tabq *q, const uint8_t *boards, const uint8_t *actions, const double *rewards,
                   const uint8_t *successors, int64_t n, void *stream);

/* Lossless key of a board (what the table stores); device [n][H*W] -> [n]. */
int sgk_board_to_key(const sgk_env *env, const uint8_t *boards, uint64_t *keys_out, int64_t n, void *stream);
/* ... and back: the board a key stands for, i.e. the reference's dict key
 * tuple(board.flatten()) (value.py:34) of a table entry.  device [n] -> [n][H*W]. */
int sgk_key_to_board(const sgk_env *env, const uint64_t *keys, uint8_t *boards_out, int64_t n, void *stream);
/* Frames per episode before the time limit ends it (safety_game max_iterations). */
int sgk_env_max_iterations(const sgk_env *env);

/* Dump table `table`: keys_out [capacity] (0 = empty slot), q_out
 * [capacity][4], corruption_out [capacity] or NULL; device pointers. */
int sgk_tabq_export(const sgk_tabq *q, int64_t table, uint64_t *keys_out, double *q_out,
                    double *corruption_out, void *stream);
/* Overwrite table `table` from the same layout (host-side merge / restore). */
int sgk_tabq_import(sgk_tabq *q, int64_t table, const uint64_t *keys, const double *qrows, void *stream);

/* The reference's dict never fills (value.py:31).  Hashed private tables
 * therefore grow on demand: sgk_tabq_act / _learn / sgk_rollout_tabq* measure
 * the fullest table when their bound on its key count nears the capacity and
 * rehash every table into twice the capacity (host-side, synchronises, rare).
 * sgk_tabq_set_auto_grow(q, 0) pins the capacity: an overflowing table then
 * drops the update, treats the state as unseen and raises SGK_EFULL at the
 * next sgk_check / *_host call.  sgk_tabq_max_fill: key count of the fullest
 * table (synchronises); sgk_tabq_grow: explicit rehash to `new_capacity`. */
int sgk_tabq_set_auto_grow(sgk_tabq *q, int enabled);
int sgk_tabq_max_fill(sgk_tabq *q, int64_t *max_fill_out, void *stream);
int sgk_tabq_grow(sgk_tabq *q, int64_t new_capacity, void *stream);

/* Replica sync for shared tables on several GPUs: every
 * GPU keeps a replica; at a sync point each exports its change since the last
 * sync, the records are all-gathered (NCCL), every replica is restored to the
 * last synced table and all deltas are applied in rank order scaled by
 * 1/n_ranks, then the result becomes the new base.  keys [capacity] (0 =
 * empty), delta [capacity][4], device pointers. */
int sgk_tabq_delta_export(sgk_tabq *q, uint64_t *keys_out, double *delta_out, void *stream);
int sgk_tabq_delta_apply(sgk_tabq *q, const uint64_t *keys, const double *delta, double scale, void *stream);
int sgk_tabq_rebase(sgk_tabq *q, void *stream);
int sgk_tabq_restore_base(sgk_tabq *q, void *stream);
/* The same sync as ONE all-reduce, for tables whose keys have a small canonical
 * index (boat 25, sokoban 36 x 36, lava 63, island 48, supervisor / whisky
 * 48 x 2; tomato has none: 0): sgk_tabq_dense_size entries of 5 doubles --
 * delta-Q[4] and a presence count.  export writes this replica's change since
 * the last sync; the caller all-reduces (sum) the array over the replicas
 * (ncclAllReduce); apply restores the base, writes
 * base + scale * sum for every key any replica holds, and makes the result the
 * new base.  Replicas end bit-identical: they apply the same reduced array. */
int64_t sgk_tabq_dense_size(const sgk_tabq *q);
int sgk_tabq_delta_export_dense(sgk_tabq *q, double *delta_out, void *stream);
int sgk_tabq_delta_apply_dense(sgk_tabq *q, const double *delta_sum, double scale, void *stream);

/* ------------------------------------------------------------ SSRL agent --
 * TabularSSQAgent (ssrl/agents.py:9-86): per-state corruption estimate C with
 * prior `c_prior`, reward scaled by 1 - C[s] in learn, and at every episode
 * end, while `budget` queries remain, query_H + learn_C over the states the
 * episode visited (loop defined in DESIGN.md; the reference ships none). */
int sgk_tabq_enable_ssrl(sgk_tabq *q, double c_prior, int64_t budget, int64_t max_episode_steps);
/* ssrl.random_warmup (ssrl/warmup.py:4-35): every environment plays
 * `n_episodes` random-policy episodes (RandomAgent, dummy.py:15-16; reset at
 * the start of each, like the reference's loop, and left un-reset after the
 * last); after each, query_H (budget -= 1) and learn_C(return - safety > 0).
 * As in the reference the warm-up never calls act_explore, so the agent's
 * history is empty and only the episode / corrupt-episode counters move
 * (ssrl/agents.py:50-82).  steps_done: device [n_envs] int64 or NULL. */
int sgk_ssrl_warmup(sgk_env *env, sgk_tabq *q, int64_t n_episodes, uint64_t t0, int64_t *steps_done, void *stream);
/* The same agent method by method, for callers that keep the episode's history
 * on the host (the N = 1 adapter): optional query_H bookkeeping (budget -= 1,
 * ssrl/agents.py:45-48), then learn_C(corrupt) over `boards` (device
 * [n_boards][H*W], the states passed to act_explore this episode, in order;
 * ssrl/agents.py:50-75), then reset_history(corrupt, increment_episode)
 * (:77-82), all for table `table`. */
int sgk_ssrl_learn_c(sgk_tabq *q, int64_t table, const uint8_t *boards, int64_t n_boards, int corrupt, int query,
                     int increment_episode, void *stream);
/* TabularSSQAgent.budget / .episodes / .corrupt_episodes per environment
 * (device [n_envs] int64 each, any may be NULL). */
int sgk_ssrl_get_counters(const sgk_tabq *q, int64_t *budget, int64_t *episodes, int64_t *corrupt_episodes, void *stream);

/* ------------------------------------------------------ fused hot kernel --
 * `n_steps` lock-steps of the whole tabq_learn body (common/learn.py:61-85)
 * inside one kernel, for every environment of `env`:
 *   act_explore -> env.step -> (--cheat: learn from hidden reward,
 *   learn.py:72-73) -> learn -> update_epsilon -> reset when done
 *   (train.py:62-70), with the episode metrics of meters.py:66-108
 *   accumulated per environment.
 * `t0` is the agent-step index of the first lock-step (continue a run by
 * passing the previous t0 + n_steps). */
int sgk_rollout_tabq(sgk_env *env, sgk_tabq *q, int64_t n_steps, uint64_t t0, int cheat, void *stream);

/* The same body run EPISODE-wise, the shape of one reference call
 * tabq_learn(agent, env, env_state, history, args) = whiler.stepbystep
 * (common/learn.py:13-24): every environment runs until it has finished
 * `max_episodes` episodes (at most `max_steps` lock-steps) and is NOT reset
 * after its last one -- like the reference, whose caller resets at the top of
 * the next episode (train.py:62-70); a finished environment is reset by its
 * next step.  Private tables only.  Outputs (device [n_envs], any may be NULL):
 * steps executed, and reward / info["hidden_reward"] (NaN = None) of the last
 * step; its actual action is read with sgk_env_actual_actions.  Meant for the
 * N = 1 adapters: with several environments the agent-step index (epsilon
 * schedule, random streams) restarts from the common `t0` of the next call. */
int sgk_rollout_tabq_episodes(sgk_env *env, sgk_tabq *q, int64_t max_episodes, int64_t max_steps, uint64_t t0,
                              int cheat, int64_t *steps_done, double *last_reward, double *last_hidden,
                              void *stream);

/* Same as sgk_rollout_tabq but with a uniform random policy and no learning
 * (RandomAgent, common/agents/dummy.py:7-16; warm-up loops). */
int sgk_rollout_random(sgk_env *env, int64_t n_steps, uint64_t t0, void *stream);

/* default_eval (common/eval.py:8-56) for every environment of `eval_env`:
 * greedy actions from the table(s) of `q` (table i for environment i, or the
 * shared table), no learning and no insertion, each environment running until
 * the first episode end at or after `eval_timesteps` steps.  Episode metrics
 * accumulate in eval_env (read them with sgk_env_totals*). */
int sgk_eval_tabq(sgk_env *eval_env, const sgk_tabq *q, int64_t eval_timesteps, uint64_t t0, void *stream);
/* ... with the two things the single-environment drop-in needs to be exact:
 * insert_on_miss -- the reference evaluates through act(), whose defaultdict
 * lookup inserts a zero row for every unseen board (value.py:31,35), so its key
 * set grows during evaluation (private tables only); episode_log -- device
 * [n_envs][log_cap][2] doubles receiving (return, performance) of every
 * evaluation episode in order, what track_metrics feeds the meters one episode
 * at a time (meters.py:76-83); NULL = not wanted. */
int sgk_eval_tabq_ex(sgk_env *eval_env, sgk_tabq *q, int64_t eval_timesteps, uint64_t t0, int insert_on_miss,
                     double *episode_log, int64_t log_cap, void *stream);

/* Check the sticky device status word (table full, replay stream dry);
 * synchronises `stream`. */
int sgk_check(sgk_env *env, sgk_tabq *q, void *stream);

/* Host-buffer form of the fused call, for callers whose data lives in host
 * memory: uploads `actions_or_null`-free state, i.e. copies the environment
 * core state [n_envs] u64 from `core_in` (host, may be NULL = keep device
 * state), runs n_steps, and copies back boards [n_envs][H*W], the 7 totals
 * and core state.  Synchronises. */
int sgk_rollout_tabq_host(sgk_env *env, sgk_tabq *q, int64_t n_steps, uint64_t t0, int cheat,
                          const uint64_t *core_in, uint64_t *core_out, uint8_t *boards_out,
                          double totals_out[SGK_N_TOTALS], void *stream);

/* -------------------------------------------------------------- deep Q --
 * DeepQAgent (common/agents/value.py:61-187) for N lock-step environments
 * sharing one Q network: MLP Linear(n_in, n_hidden)-ReLU-[Linear-ReLU] x
 * (n_layers-1)-Linear(n_hidden, n_actions) (value.py:148-158), a target
 * network, Adam(amsgrad=True) (value.py:87) and the replay buffer
 * (common/utils/contain.py:8-22) as a ring of packed uint8 transitions in HBM.
 * n_in is H*W*C of the board (the reference multiplies only two of the three
 * dims, value.py:66-67 -- a shape bug that makes it unrunnable). */
typedef struct sgk_dqn sgk_dqn;
int sgk_dqn_create(const sgk_env *env, int n_layers, int n_hidden, int64_t replay_capacity, int64_t batch_size,
                   uint64_t seed, sgk_dqn **out);
int sgk_dqn_destroy(sgk_dqn *d);
/* args.lr, .discount, .epsilon, .epsilon_anneal (value.py:72-79), args.sync_every
 * (learn.py:55) and whether to reproduce the reference loss, which broadcasts
 * Qs[B,1] against expected_Qs[B] to a B x B mean (value.py:119-123). */
int sgk_dqn_configure(sgk_dqn *d, double lr, double discount, double epsilon, int64_t epsilon_anneal,
                      int64_t sync_every, int reference_bxb_loss);
int64_t sgk_dqn_param_count(const sgk_dqn *d);
int64_t sgk_dqn_replay_count(const sgk_dqn *d);
/* Flat parameters in torch order (weight, bias per Linear); which: 0 = Q,
 * 1 = target_Q; device float32 [param_count]. */
int sgk_dqn_get_params(const sgk_dqn *d, int which, float *out, void *stream);
int sgk_dqn_set_params(sgk_dqn *d, int which, const float *in, void *stream);
/* Gradients of the last learn step before clipping (what the reference logs as
 * histograms with --log-gradients, value.py:129-133); same flat layout. */
int sgk_dqn_get_grads(const sgk_dqn *d, float *out, void *stream);
/* sync_target_Q (value.py:138-140) */
int sgk_dqn_sync_target(sgk_dqn *d, void *stream);
/* scores of DeepQAgent.act (value.py:89-92): boards [n][H*W] u8 -> q_out [n][4] f32 */
int sgk_dqn_qvalues(sgk_dqn *d, int which, const uint8_t *boards, int64_t n, float *q_out, void *stream);
/* ReplayBuffer.add for n transitions (contain.py:15-17) */
int sgk_dqn_replay_add(sgk_dqn *d, const uint8_t *s, const uint8_t *a, const double *r, const uint8_t *s2,
                       const uint8_t *term, int64_t n, void *stream);
/* Rows [first, first + n) of the ring, i.e. ReplayBuffer.buffer entries in
 * insertion order modulo capacity (contain.py:8-17): device arrays, any may be
 * NULL.  Inspection / checkpointing; not on the hot path. */
int sgk_dqn_replay_get(const sgk_dqn *d, int64_t first, int64_t n, uint8_t *s, uint8_t *a, float *r,
                       uint8_t *s2, uint8_t *term, void *stream);
/* DeepQAgent.learn after replay.add (value.py:115-136): sample batch_size
 * transitions with replacement, loss, backward, clip_grad_norm_(10), Adam.
 * loss_out (device float[3], may be NULL) = loss, gradient norm, clip factor. */
int sgk_dqn_learn(sgk_dqn *d, uint64_t step, float *loss_out, void *stream);
/* The same optimiser step on an explicit batch (parity tests against torch). */
int sgk_dqn_learn_batch(sgk_dqn *d, const uint8_t *s, const uint8_t *a, const double *r, const uint8_t *s2,
                        const uint8_t *term, int64_t n, float *loss_out, void *stream);
int sgk_dqn_last_scalars(const sgk_dqn *d, float *out3, void *stream);
/* How the network math runs.  The reference's default architecture (n_layers 2,
 * n_hidden <= 100, value.py:148-158) runs on the 5th-generation tensor cores:
 * every forward pass (acting, online and target networks in learn) as one fused
 * tcgen05 kernel with activations resident in tensor memory, the backward pass
 * as an error-chain kernel plus sample-reduction weight-gradient kernels.
 *   mode 3 (default where supported)  3xTF32 forward: every fp32 operand split
 *           into two TF32 numbers, three MMA passes, fp32 accumulation -- Q
 *           values within 1e-5 of torch fp32 (the tolerance north_star states);
 *           backward in single-pass TF32
 *   mode 1  single-pass TF32 everywhere (about 1e-3)
 *   mode 0  fp32 FFMA kernels (the parity reference for gradients / optimiser;
 *           the only mode for other architectures) */
int sgk_dqn_set_tensor_cores(sgk_dqn *d, int mode);
int sgk_dqn_get_tensor_cores(const sgk_dqn *d);
/* n_steps lock-steps of the dqn_learn body (common/learn.py:29-58) for every
 * environment: act_explore, env.step, replay.add, learn, update_epsilon,
 * target sync every sync_every steps, reset when done.  `mode` is a bit set:
 * without SGK_DQN_LEARN the random-policy warm-up that only fills the ring
 * runs (common/warmup.py:8-23); SGK_DQN_CHEAT is args.cheat (learn.py:39-47):
 * the transition stores info["hidden_reward"] (None -> 0) as its reward and
 * the action the environment really executed ("actual_actions"). */
#define SGK_DQN_LEARN 1
#define SGK_DQN_CHEAT 2
int sgk_rollout_dqn(sgk_env *env, sgk_dqn *d, int64_t n_steps, uint64_t t0, int mode, void *stream);

/* get_discounted_returns (common/agents/policy_base.py:179-186) for a block of
 * n_steps lock-steps collected from `env`: reward / done are device
 * [n_steps][n_envs], frame0 [n_envs] the in-episode index of row 0; returns
 * float32 [n_steps][n_envs].  Used by the batched rollout collector that feeds
 * policy-gradient agents (gather_rollout, policy_base.py:133-177). */
int sgk_discounted_returns(const sgk_env *env, const double *reward, const uint8_t *done, const int32_t *frame0,
                           int64_t n_steps, double discount, float *returns_out, void *stream);

/* Raw environment state words, device [n_envs] (checkpoint / e2e path). */
int sgk_env_get_core(const sgk_env *env, uint64_t *core_out, void *stream);
int sgk_env_set_core(sgk_env *env, const uint64_t *core_in, void *stream);

/* Enable the per-step trace hash in the fused kernels (test mode). */
int sgk_env_set_trace(sgk_env *env, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* SGK_H */
