// TEST INFRASTRUCTURE: host (g++) build of the step kernel's per-phase logic
// (crowdnav_prediction_attngraph_b200/csrc/cn_env_core.cuh) so the environment arithmetic can be checked
// against the oracle / golden vectors without a GPU.  The phase ORDER below mirrors
// cn_env_step_kernel in cn_env_kernels.cu; thread barriers become plain loops over humans.
// This is NOT a CPU fallback of the product: nothing under crowdnav_prediction_attngraph_b200/ loads it.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

struct float4 { float x, y, z, w; };
#include "../../crowdnav_prediction_attngraph_b200/csrc/cn_env_core.cuh"
#include "../../include/crowdnav_b200.h"

struct Harness {
  CnParams p;
  CnState g;
  std::map<std::string, std::pair<void*, size_t>> fields;
  std::vector<void*> allocs;
};

template <class T>
static void halloc(Harness* hn, const char* name, T** ptr, size_t count) {
  *ptr = static_cast<T*>(calloc(count ? count : 1, sizeof(T)));
  hn->allocs.push_back(*ptr);
  hn->fields[name] = {(void*)*ptr, count * sizeof(T)};
}

template <int MAXH>
static void run(Harness* hn, const float* action, const cn_obs_ptrs* o, const cn_step_ptrs* r, int mode) {
  const CnParams& p = hn->p;
  CnState& g = hn->g;
  const int H = p.H;
  CnObs ob{o->robot_node, o->temporal_edges, o->spatial_edges, o->detected_human_num, o->visible_masks};
  CnStepOut out;
  memset(&out, 0, sizeof(out));
  if (r) out = CnStepOut{r->reward, r->done, r->info, r->info_aux, r->ep_ret, r->ep_len, r->not_done};
  std::vector<double> d(12 * H);
  std::vector<float> f(6 * H);
  std::vector<uint8_t> u(H);
  std::vector<float4> lines((size_t)H * H);
  std::vector<float> rows((size_t)H * 16);
  std::vector<float4> projbuf(MAXH);
  for (int e = 0; e < p.N; ++e) {
    CnEnvSh s;
    s.px = d.data(); s.py = s.px + H; s.gx = s.py + H; s.gy = s.gx + H; s.rad = s.gy + H; s.vpref = s.rad + H;
    s.t0 = s.vpref + H; s.t1 = s.t0 + H;
    s.wx = s.t1 + H; s.wy = s.wx + H; s.nwx = s.wy + H; s.nwy = s.nwx + H;      // social-force humans only
    s.vx = f.data(); s.vy = s.vx + H; s.fx = s.vy + H; s.fy = s.fx + H; s.nvx = s.fy + H; s.nvy = s.nvx + H;
    s.visr = u.data();
    s.lean = 0;
    const CnCoop co = {0, 1, nullptr, nullptr};
    uint32_t* prep_key = g.prep_mt + (size_t)e * 624;
    if (mode == 1) {
      // full reset = prepare (event kernel, forced) -> install + first observation (step kernel, mode 1)
      cn_prepare_env(p, g, s, e, prep_key, co);
      s.done = 1; s.info = 0; s.reward = 0.0; s.reset_flag = 0; s.nvis = 0; s.goal_flag = 0; s.hn = 0;
      for (int h = H - 1; h >= 0; --h) cn_install_env(p, g, s, e, h);
    } else {
      for (int h = H - 1; h >= 0; --h) cn_phase_load(p, g, s, e, h, action);
      const int hn = s.hn;                       // live humans (slots [hn, H) are empty)
      if (p.social_force) for (int h = 0; h < hn; ++h) cn_sf_action(p, g, s, e, h);
      for (int h = 0; h < hn && !p.social_force; ++h) {
        // single-lane "warp": the cooperative solver degenerates to the sequential RVO2 order
        CnWarpLines W; W.smem0 = lines.data() + (size_t)h * H; W.stride = 1; W.cap = 3;    // exercise both tiers
        W.ovf0 = lines.data() + (size_t)h * H + 3; W.ovf_stride = 0;
        CnLineStore proj; proj.base = projbuf.data(); proj.stride = 1; proj.cap = MAXH; proj.ovf = nullptr;
        int nl = 0, fail = -1; float vmax = 0; CnF2 pref = f2(0, 0), result = f2(0, 0);
        cn_orca_build<MAXH>(p, g, s, e, h, W.of(0), nl, vmax, pref);
        cn_orca_lp2_warp(co, W, nl, vmax, pref, result, fail);
        cn_orca_lp3_warp(co, W, nl, vmax, proj, result, fail);
        cn_orca_finish(p, g, s, e, h, result, nl, fail);
      }
      if (p.test_phase) {
        // ground-truth look-ahead (phase 'test'): lookahead_steps nested ORCA solves of every human on a
        // scratch copy of the joint state; mirrors the look-ahead loop of cn_env_step_kernel
        std::vector<double> sx(s.px, s.px + H), sy(s.py, s.py + H), lx(sx), ly(sy);
        std::vector<float> svx(s.vx, s.vx + H), svy(s.vy, s.vy + H), lvx(svx), lvy(svy);
        std::vector<CnLookahead> la(H);
        std::vector<CnF2> res(H);
        std::vector<int> nls(H), fails(H);
        for (int h = 0; h < hn; ++h) { la[h].min_rd = INFINITY; la[h].pen = 0.0; }
        for (int t = 1; t <= p.lookahead_steps; ++t) {
          for (int h = 0; h < hn; ++h) {
            s.px[h] = lx[h]; s.py[h] = ly[h]; s.fx[h] = (float)lx[h]; s.fy[h] = (float)ly[h];
            s.vx[h] = lvx[h]; s.vy[h] = lvy[h];
          }
          for (int h = 0; h < hn; ++h) {
            CnWarpLines W; W.smem0 = lines.data() + (size_t)h * H; W.stride = 1; W.cap = 3;
            W.ovf0 = lines.data() + (size_t)h * H + 3; W.ovf_stride = 0;
            CnLineStore proj; proj.base = projbuf.data(); proj.stride = 1; proj.cap = MAXH; proj.ovf = nullptr;
            int nl = 0, fail = -1; float vmax = 0; CnF2 pref = f2(0, 0), result = f2(0, 0);
            cn_orca_build<MAXH>(p, g, s, e, h, W.of(0), nl, vmax, pref, false);
            cn_orca_lp2_warp(co, W, nl, vmax, pref, result, fail);
            cn_orca_lp3_warp(co, W, nl, vmax, proj, result, fail);
            res[h] = result; nls[h] = nl; fails[h] = fail;
          }
          for (int h = 0; h < hn; ++h) {
            lx[h] = lx[h] + (double)res[h].x * p.time_step; ly[h] = ly[h] + (double)res[h].y * p.time_step;
            lvx[h] = res[h].x; lvy[h] = res[h].y;
            if (t % p.pred_interval == 0)
              cn_lookahead_accumulate(p, s, g.vis[cn_idx(p, e, h)] != 0, lx[h], ly[h], t / p.pred_interval, la[h]);
            if (t == p.lookahead_steps) cn_orca_diag(p, g, e, h, res[h], nls[h], fails[h]);
          }
        }
        for (int h = 0; h < hn; ++h) {
          s.px[h] = sx[h]; s.py[h] = sy[h]; s.fx[h] = (float)sx[h]; s.fy[h] = (float)sy[h];
          s.vx[h] = svx[h]; s.vy[h] = svy[h];
          s.t0[h] = la[h].min_rd; s.t1[h] = la[h].pen;
        }
      }
      cn_phase_reward(p, g, s, e, out);
      if (s.done) { for (int h = H - 1; h >= 0; --h) cn_install_env(p, g, s, e, h); }   // prepared next episode
      else { for (int h = 0; h < hn; ++h) cn_phase_integrate(p, s, h); }
      if (cn_add_remove_due(p, g, s, e)) cn_phase_add_remove(p, g, s, e);
    }
    for (int h = 0; h < H; ++h) cn_phase_obs_a<16>(p, g, s, e, h, rows.data() + (size_t)h * 16);
    for (int h = 0; h < H; ++h) cn_phase_obs_b(p, g, s, e, h, rows.data() + (size_t)h * 16, ob);
    for (int h = 0; h < H; ++h) cn_phase_obs_c(p, s, e, h, ob);
    const int evt = cn_event_flag(p, g, s, e);
    // event kernel: goal dynamics (evt 1) or preparation of the next episode (evt 2)
    if (evt == 1) cn_phase_goals(p, g, s, e, g.mt + (size_t)e * 624, co);
    for (int h = 0; h < H; ++h) cn_phase_store(p, g, s, e, h);
    if (evt == 2) cn_prepare_env(p, g, s, e, prep_key, co);
  }
}

extern "C" {

void* harness_create(const cn_config* cfg) {
  Harness* hn = new Harness();
  CnParams& p = hn->p;
  memset(&p, 0, sizeof(p));
  p.hbase = cfg->human_num; p.hrange = cfg->human_num_range;
  p.N = cfg->num_envs; p.H = cfg->human_num + cfg->human_num_range; p.P = cfg->predict_steps;
  p.const_vel = cfg->const_vel ? 1 : 0;
  p.W = p.const_vel ? 2 * (p.P + 1) : 2;
  p.randomize = cfg->randomize_attributes; p.goal_changing = cfg->random_goal_changing;
  p.end_goal_changing = cfg->end_goal_changing; p.sort_humans = cfg->sort_humans;
  p.nenv_total = cfg->nenv_total; p.seed_base = (uint32_t)(cfg->seed + cfg->rank_offset);
  p.time_step = cfg->time_step; p.time_limit = cfg->time_limit;
  p.pred_dt = cfg->time_step * (double)(int)floor(cfg->pred_timestep / cfg->time_step);
  cn_fill_phase(p, cfg->phase, cfg->val_size, cfg->test_size);
  p.circle_radius = cfg->circle_radius; p.arena_size = cfg->arena_size;
  p.discomfort_dist = cfg->discomfort_dist; p.discomfort_penalty_factor = cfg->discomfort_penalty_factor;
  p.success_reward = cfg->success_reward; p.collision_penalty = cfg->collision_penalty;
  p.human_radius = cfg->human_radius; p.human_vpref = cfg->human_v_pref;
  p.robot_radius = cfg->robot_radius; p.robot_vpref = cfg->robot_v_pref; p.sensor_range = cfg->sensor_range;
  p.human_fov = CN_PI * cfg->human_fov; p.robot_fov = CN_PI * cfg->robot_fov;
  p.goal_change_chance = cfg->goal_change_chance;
  p.orca_safety_space = cfg->orca_safety_space; p.orca_neighbor_dist = cfg->orca_neighbor_dist;
  p.orca_time_horizon = (float)cfg->orca_time_horizon;
  p.social_force = cfg->human_policy == 1 ? 1 : 0; p.sf_A = cfg->sf_A; p.sf_B = cfg->sf_B; p.sf_KI = cfg->sf_KI;
  const size_t N = p.N, NH = N * p.H;
  CnState& g = hn->g;
#define A(field, count) halloc(hn, #field, &g.field, (count))
  A(rpx, N); A(rpy, N); A(rgx, N); A(rgy, N); A(rvx, N); A(rvy, N); A(potential, N); A(fut_pen, N);
  A(nd_global, N); A(ep_ret, N); A(ep_len, N); A(step_count, N); A(case_counter, N); A(seed_off, N);
  A(hpx, NH); A(hpy, NH); A(hgx, NH); A(hgy, NH); A(hrad, NH); A(hvpref, NH); A(hvx, NH); A(hvy, NH);
  A(bpx, NH); A(bpy, NH); A(bvx, NH); A(bvy, NH); A(brad, NH); A(vis, NH);
  A(sim_exists, NH); A(sim_nd, NH); A(sim_rself, NH); A(sim_vmax, NH); A(sim_rother, NH * p.H);
  A(mt, N * 624); A(mt_pos, N);
  A(prep_robot, N * 4); A(prep_hpx, NH); A(prep_hpy, NH); A(prep_hrad, NH); A(prep_hvpref, NH); A(prep_nd, N);
  A(prep_mt, N * 624); A(prep_mt_pos, N);
  A(last_hvx, NH); A(last_hvy, NH); A(orca_nlines, NH); A(orca_fail, NH); A(evt, N); A(spawn_overflow, N); A(lp_cost, N); A(hn, N); A(prep_hn, N); A(sim_n, NH); A(hwx, NH); A(hwy, NH);
#undef A
  for (size_t e = 0; e < N; ++e) { g.nd_global[e] = cfg->orca_neighbor_dist; g.seed_off[e] = (int32_t)e; }
  return hn;
}

void harness_destroy(void* h) {
  Harness* hn = static_cast<Harness*>(h);
  for (void* q : hn->allocs) free(q);
  delete hn;
}

static void dispatch(Harness* hn, const float* a, const cn_obs_ptrs* o, const cn_step_ptrs* r, int mode) {
  if (hn->p.H <= 32) run<32>(hn, a, o, r, mode);
  else if (hn->p.H <= 64) run<64>(hn, a, o, r, mode);
  else run<128>(hn, a, o, r, mode);
}

void harness_reset(void* h, const cn_obs_ptrs* o) { dispatch(static_cast<Harness*>(h), nullptr, o, nullptr, 1); }
void harness_step(void* h, const float* action, const cn_obs_ptrs* o, const cn_step_ptrs* r) {
  dispatch(static_cast<Harness*>(h), action, o, r, 0);
}
size_t harness_state_bytes(void* h, const char* name) {
  Harness* hn = static_cast<Harness*>(h);
  auto it = hn->fields.find(name);
  return it == hn->fields.end() ? 0 : it->second.second;
}
int harness_state_copy(void* h, const char* name, void* buf, size_t bytes, int dir) {
  Harness* hn = static_cast<Harness*>(h);
  auto it = hn->fields.find(name);
  if (it == hn->fields.end() || it->second.second != bytes) return 1;
  if (dir) memcpy(it->second.first, buf, bytes); else memcpy(buf, it->second.first, bytes);
  return 0;
}

// MT19937 unit-test hooks
void harness_rng_doubles(uint32_t seed, int n, double* out) {
  uint32_t key[624];
  CnRng r; r.key = key; r.pos = 624;
  const CnCoop co = {0, 1, nullptr, nullptr};
  cn_rng_seed(r, seed, co);
  for (int i = 0; i < n; ++i) out[i] = cn_rng_double(r, co);
}

}  // extern "C"
