"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of ONE reference crowd-navigation environment (holonomic robot, ORCA
humans), written as flat arrays + scalar fp64 arithmetic instead of the reference's
Agent/Human/Robot object graph.  It follows, step for step, the order of operations of

  crowd_sim/envs/crowd_sim_pred.py:100-213   CrowdSimPred.step
  crowd_sim/envs/crowd_sim_var_num.py:303-363 CrowdSimVarNum.reset
  crowd_sim/envs/crowd_sim_var_num.py:465-561 calc_reward (+ crowd_sim_pred.py:216-233)
  crowd_sim/envs/crowd_sim_pred.py:62-97      generate_ob  (VarNum: crowd_sim_var_num.py:233-279)
  crowd_sim/envs/crowd_sim.py:243-273,513-572,680-703  belief update / visibility / human actions
  crowd_sim/envs/crowd_sim.py:415-450         update_human_goals_randomly
  crowd_sim/envs/crowd_sim_var_num.py:116-146 generate_circle_crossing_human
  crowd_nav/policy/orca.py:64-117             ORCA.predict (one persistent rvo2 sim per human)
  crowd_sim/envs/utils/agent.py:20-23,44-50   per-Agent RNG draws when randomize_attributes
  rl/networks/shmem_vec_env.py:138-142        auto-reset on done, float32 observation buffers

It is pinned against the UNMODIFIED reference run in this container (tools/make_golden.py →
tests/golden/env_*.npz, tests/test_oracle_golden.py).  The ORCA arithmetic itself lives in
oracle/rvo2_ref.cpp ("parity unpinned": the reference vendors no rvo2 source or version).

Scope: sim.human_num_range == 0, holonomic kinematics, ORCA humans, robot.visible == False
(the reference defaults, and every BASELINE.json config).
"""
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.join(_HERE, "shims") not in sys.path:
    sys.path.insert(0, os.path.join(_HERE, "shims"))
import rvo2  # noqa: E402  (oracle/shims/rvo2.py -> oracle/rvo2_ref.cpp)

INFO_NOTHING, INFO_TIMEOUT, INFO_COLLISION, INFO_REACHGOAL, INFO_DANGER = 0, 1, 2, 3, 4
INFO_NAMES = {0: "", 1: "Timeout", 2: "Collision", 3: "Reaching goal", 4: "Too close"}


class EnvConfig(object):
    """Flat snapshot of the reference `Config` fields the hot path reads
    (crowd_nav/configs/config.py:16-120, arguments.py:206)."""

    def __init__(self, **kw):
        self.human_num = 20
        self.human_num_range = 0             # sim.human_num_range: humans join / leave every 5 s
        self.human_policy = "orca"           # humans.policy: 'orca' | 'social_force' (crowd_nav/policy/social_force.py)
        self.sf_A, self.sf_B, self.sf_KI = 2.0, 1.0, 1.0     # config.sf
        self.predict_steps = 5
        self.predict_method = "const_vel"      # 'const_vel' | 'none' (CrowdSimVarNum-v0 obs)
        self.time_limit = 50.0
        self.time_step = 0.25
        self.pred_timestep = 0.25              # config.data.pred_timestep
        self.randomize_attributes = False
        self.random_goal_changing = False
        self.goal_change_chance = 0.5
        self.end_goal_changing = True
        self.circle_radius = 6 * np.sqrt(2)
        self.arena_size = 6
        self.success_reward = 10
        self.collision_penalty = -20
        self.discomfort_dist = 0.25
        self.discomfort_penalty_factor = 10
        self.human_radius = 0.3                # config.humans.radius
        self.human_v_pref = 1
        self.human_fov = 2.0                   # x pi
        self.robot_radius = 0.3
        self.robot_v_pref = 1
        self.robot_fov = 2.0                   # x pi
        self.sensor_range = 5
        self.orca_neighbor_dist = 10
        self.orca_safety_space = 0.15
        self.orca_time_horizon = 5
        self.sort_humans = True
        self.val_size = 100
        self.test_size = 500
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown EnvConfig field %r" % k)
            setattr(self, k, v)

    @classmethod
    def from_reference(cls, config):
        """Snapshot a reference `crowd_nav.configs.config.Config` object."""
        assert config.action_space.kinematics == "holonomic"
        assert config.humans.policy in ("orca", "social_force") and not config.robot.visible
        pm = config.sim.predict_method
        return cls(
            human_num=config.sim.human_num, human_num_range=config.sim.human_num_range, predict_steps=config.sim.predict_steps,
            human_policy=config.humans.policy, sf_A=config.sf.A, sf_B=config.sf.B, sf_KI=config.sf.KI,
            predict_method=pm, time_limit=config.env.time_limit, time_step=config.env.time_step,
            pred_timestep=config.data.pred_timestep,
            randomize_attributes=config.env.randomize_attributes,
            random_goal_changing=config.humans.random_goal_changing,
            goal_change_chance=config.humans.goal_change_chance,
            end_goal_changing=config.humans.end_goal_changing,
            circle_radius=config.sim.circle_radius, arena_size=config.sim.arena_size,
            success_reward=config.reward.success_reward,
            collision_penalty=config.reward.collision_penalty,
            discomfort_dist=config.reward.discomfort_dist,
            discomfort_penalty_factor=config.reward.discomfort_penalty_factor,
            human_radius=config.humans.radius, human_v_pref=config.humans.v_pref,
            human_fov=config.humans.FOV, robot_radius=config.robot.radius,
            robot_v_pref=config.robot.v_pref, robot_fov=config.robot.FOV,
            sensor_range=config.robot.sensor_range,
            orca_neighbor_dist=config.orca.neighbor_dist,
            orca_safety_space=config.orca.safety_space,
            orca_time_horizon=config.orca.time_horizon,
            sort_humans=config.args.sort_humans, val_size=config.env.val_size,
            test_size=config.env.test_size)


def _norm2(x, y):
    # same call form as the reference's `norm((a, b))` (numpy: sqrt(x.dot(x)))
    return np.linalg.norm((x, y))


class CrowdEnvOracle(object):
    """One environment.  `this_seed`/`nenv`/`phase` as set by rl/networks/envs.py:51-58."""

    def __init__(self, cfg, this_seed, nenv, phase="train"):
        if phase not in ("train", "val", "test"):
            raise ValueError("phase must be 'train', 'val' or 'test'")
        self.cfg = cfg
        self.H = cfg.human_num               # CURRENT number of humans (changes when human_num_range > 0)
        self.Hmax = cfg.human_num + cfg.human_num_range
        self.Hmin = cfg.human_num - cfg.human_num_range
        assert cfg.human_num > cfg.human_num_range
        self.observed_human_ids = []
        self.P = cfg.predict_steps
        self.this_seed = this_seed
        self.nenv = nenv
        self.phase = phase
        self.case_counter = {"train": 0, "val": 0, "test": 0}
        u32 = int(np.iinfo(np.uint32).max)
        self.case_size = {"train": u32 - 2000, "val": cfg.val_size, "test": cfg.test_size}
        self.pred_interval = int(cfg.pred_timestep // cfg.time_step)
        self.rng = np.random.RandomState(0)
        # config.orca.neighbor_dist: process-global in the reference (agent.py:21-22)
        self.nd_global = cfg.orca_neighbor_dist
        self.last_human_states = np.zeros((self.H, 5))
        self.human_future_traj = None
        self.ep_ret = 0.0   # bench.Monitor bookkeeping
        self.ep_len = 0
        self.W = 2 * (self.P + 1) if cfg.predict_method == "const_vel" else 2

    # ---------------------------------------------------------------- RNG-consuming helpers
    def _new_human_attrs(self):
        """Human(...) construction: agent.py:20-23 then sample_random_attributes (agent.py:44-50)."""
        c = self.cfg
        v_pref, radius = c.human_v_pref, c.human_radius
        if c.randomize_attributes:
            self.nd_global = self.rng.uniform(5, 10)
            v_pref = self.rng.uniform(0.5, 1.5)
            radius = self.rng.uniform(0.3, 0.5)
        return v_pref, radius

    def _circle_crossing_human(self, skip=None):
        """crowd_sim_var_num.py:116-146.  Checks against robot + every human currently in the
        list (for a respawn the list still contains the human being replaced)."""
        c = self.cfg
        v_pref, radius = self._new_human_attrs()
        while True:
            angle = self.rng.random_sample() * np.pi * 2
            px_noise = self.rng.uniform(0, 1) * 2
            py_noise = self.rng.uniform(0, 1) * 2
            px = c.circle_radius * np.cos(angle) + px_noise
            py = c.circle_radius * np.sin(angle) + py_noise
            collide = False
            n_present = len(self.hpx)
            for k in range(-1, n_present):
                if k < 0:
                    ax, ay, agx, agy, ar = self.rpx, self.rpy, self.rgx, self.rgy, c.robot_radius
                else:
                    ax, ay, agx, agy, ar = self.hpx[k], self.hpy[k], self.hgx[k], self.hgy[k], self.hrad[k]
                min_dist = radius + ar + c.discomfort_dist
                if _norm2(px - ax, py - ay) < min_dist or _norm2(px - agx, py - agy) < min_dist:
                    collide = True
                    break
            if not collide:
                break
        return px, py, v_pref, radius

    # ---------------------------------------------------------------- reset
    def reset(self):
        c = self.cfg
        phase = self.phase
        self.global_time = 0
        self.step_counter = 0
        offset = {"train": 2000, "val": 0, "test": 1000}[phase]
        self.rand_seed = offset + self.case_counter[phase] + self.this_seed
        self.rng.seed(self.rand_seed)
        # robot (crowd_sim_var_num.py:95-101)
        while True:
            px, py, gx, gy = self.rng.uniform(-c.arena_size, c.arena_size, 4)
            if np.linalg.norm([px - gx, py - gy]) >= 8:
                break
        self.rpx, self.rpy, self.rgx, self.rgy = px, py, gx, gy
        self.rvx, self.rvy = 0, 0
        self.rtheta = np.pi / 2
        # humans: randint(lo, hi) with hi - lo == 1 consumes no draw (human_num_range == 0)
        self.observed_human_ids = []
        if c.human_num_range > 0:                # crowd_sim_var_num.py:103-104
            self.H = int(self.rng.randint(low=c.human_num - c.human_num_range, high=c.human_num + c.human_num_range + 1))
        self.hpx, self.hpy, self.hvx, self.hvy = [], [], [], []
        self.hgx, self.hgy, self.hrad, self.hvpref = [], [], [], []
        self.sims = []
        for _ in range(self.H):
            px, py, v_pref, radius = self._circle_crossing_human()
            self.hpx.append(px); self.hpy.append(py); self.hvx.append(0); self.hvy.append(0)
            self.hgx.append(-px); self.hgy.append(-py); self.hrad.append(radius); self.hvpref.append(v_pref)
            self.sims.append(None)
        self.last_human_states = np.zeros((self.H, 5))
        self.case_counter[phase] = (self.case_counter[phase] + int(self.nenv)) % self.case_size[phase]
        self.potential = -abs(np.linalg.norm(np.array([self.rgx, self.rgy]) - np.array([self.rpx, self.rpy])))
        self.ep_ret, self.ep_len = 0.0, 0
        return self._generate_ob(reset=True)

    # ---------------------------------------------------------------- visibility
    def _in_fov(self, x1, y1, vx1, vy1, x2, y2, fov):
        """crowd_sim.py:513-541 (holonomic heading = atan2(vy, vx))."""
        if fov >= 2 * np.pi:
            # offset = arccos(.) is in [0, pi] <= fov/2 unless it is NaN, which only happens for a
            # zero v_12 (coincident centres): exact shortcut for the default FOV = 2 pi.
            return not (x1 == x2 and y1 == y2)
        real_theta = np.arctan2(vy1, vx1)
        v_fov = [np.cos(real_theta), np.sin(real_theta)]
        v_12 = [x2 - x1, y2 - y1]
        with np.errstate(invalid="ignore", divide="ignore"):
            v_fov = v_fov / np.linalg.norm(v_fov)
            v_12 = v_12 / np.linalg.norm(v_12)
            offset = np.arccos(np.clip(np.dot(v_fov, v_12), a_min=-1, a_max=1))
        return bool(np.abs(offset) <= fov / 2)

    def _robot_sees(self, j):
        c = self.cfg
        in_fov = self._in_fov(self.rpx, self.rpy, self.rvx, self.rvy, self.hpx[j], self.hpy[j],
                              np.pi * c.robot_fov)
        dist = np.linalg.norm([self.rpx - self.hpx[j], self.rpy - self.hpy[j]]) - c.robot_radius - self.hrad[j]
        return in_fov and bool(dist <= c.sensor_range)

    # ---------------------------------------------------------------- observation
    def _generate_ob(self, reset):
        c, H, P = self.cfg, self.H, self.P
        vis = [self._robot_sees(j) for j in range(H)]
        self.human_visibility = vis
        num_vis = sum(vis)
        robot_node = [self.rpx, self.rpy, c.robot_radius, self.rgx, self.rgy, c.robot_v_pref, self.rtheta]
        self.prev_human_pos = self.last_human_states.copy()
        # belief update (crowd_sim.py:243-273)
        for i in range(H):
            if vis[i]:
                self.last_human_states[i, :] = [self.hpx[i], self.hpy[i], self.hvx[i], self.hvy[i], self.hrad[i]]
            elif reset:
                self.last_human_states[i, :] = [15., 15., 0., 0., 0.3]
            else:
                px, py, vx, vy, r = self.last_human_states[i, :]
                self.last_human_states[i, :] = [px + vx * c.time_step, py + vy * c.time_step, vx, vy, r]
        temporal = np.array([self.rvx, self.rvy])
        if c.predict_method == "const_vel":
            # calc_human_future_traj('const_vel') (crowd_sim_var_num.py:152-228)
            traj = np.zeros((P + 1, H, 4))
            for i in range(H):
                traj[0, i, 0], traj[0, i, 1] = self.hpx[i], self.hpy[i]
            traj[0, :, 2:4] = self.prev_human_pos[:, 2:4]
            traj = np.tile(traj[0].reshape(1, H, 4), (P + 1, 1, 1))
            t = (np.arange(0, P + 1, dtype=float).reshape((P + 1, 1, 1)) * c.time_step * self.pred_interval)
            traj[:, :, :2] = traj[:, :, :2] + t * traj[:, :, 2:]
            inv = np.logical_not(vis)
            traj[:, inv, :2] = 15
            traj[:, inv, 2:] = 0
            self.human_future_traj = traj
            spatial = np.ones((self.Hmax, 2 * (P + 1))) * np.inf        # storage for max_human_num humans
            pred_pos = np.transpose(traj[:, :, :2], (1, 0, 2)) - np.array([self.rpx, self.rpy])
            spatial[:H][np.array(vis, dtype=bool)] = pred_pos.reshape((H, -1))[np.array(vis, dtype=bool)]
            if c.sort_humans:
                spatial = np.array(sorted(spatial, key=lambda x: np.linalg.norm(x[:2])))
            spatial[np.isinf(spatial)] = 15
            vmask = None
        else:
            # CrowdSimVarNum.generate_ob (crowd_sim_var_num.py:233-279)
            spatial = np.ones((self.Hmax, 2)) * np.inf
            for i in range(H):
                if vis[i]:
                    spatial[i, :] = [self.last_human_states[i, 0] - self.rpx, self.last_human_states[i, 1] - self.rpy]
            vmask = np.zeros(self.Hmax, dtype=bool)
            if c.sort_humans:
                spatial = np.array(sorted(spatial, key=lambda x: np.linalg.norm(x)))
                if num_vis > 0:
                    vmask[:num_vis] = True
            else:
                vmask[:H] = vis
            spatial[np.isinf(spatial)] = 15
            self.observed_human_ids = np.where(vis)[0]                 # only CrowdSimVarNum.generate_ob updates it
        ob = {
            "robot_node": np.asarray(robot_node, dtype=np.float32).reshape(1, 7),
            "temporal_edges": np.asarray(temporal, dtype=np.float32).reshape(1, 2),
            "spatial_edges": np.asarray(spatial, dtype=np.float32),
            "detected_human_num": np.asarray([num_vis if num_vis > 0 else 1], dtype=np.float32),
        }
        if vmask is not None:
            ob["visible_masks"] = vmask
        return ob

    # ---------------------------------------------------------------- human actions (ORCA)
    def _human_actions(self):
        """get_human_actions (crowd_sim.py:680-703): ORCA on the current state, FOV-filtered neighbours."""
        return self._orca_actions(self.hpx, self.hpy, self.hvx, self.hvy, use_fov=True)

    def _orca_actions(self, px, py, vx, vy, use_fov):
        """One ORCA solve per human on the joint state (px, py, vx, vy) through its PERSISTENT rvo2 simulator
        (crowd_nav/policy/orca.py:64-117).  use_fov=False is the ground-truth look-ahead's
        act_joint_state (crowd_sim_var_num.py:183-196): every other human is passed as is."""
        c, H = self.cfg, self.H
        fov = np.pi * c.human_fov
        acts = []
        self.last_orca_diag = []
        if c.human_policy == "social_force":
            return self._social_force_actions(px, py, vx, vy, use_fov)
        for i in range(H):
            others = []
            for j in range(H):
                if j == i:
                    continue
                # humans have no range limit; FOV test only (always true for FOV = 2 pi unless NaN)
                if (not use_fov) or self._in_fov(px[i], py[i], vx[i], vy[i], px[j], py[j], fov):
                    others.append((px[j], py[j], vx[j], vy[j], self.hrad[j]))
                else:
                    others.append((7, 7, 0, 0, 0.3))   # dummy_human (crowd_sim.py:130-133)
            sim = self.sims[i]
            if sim is not None and sim.getNumAgents() != len(others) + 1:     # orca.py:80-82: humans joined / left
                sim = self.sims[i] = None
            if sim is None:
                params = (self.nd_global, len(others), c.orca_time_horizon, c.orca_time_horizon)
                sim = rvo2.PyRVOSimulator(c.time_step, *params, self.hrad[i], 1)
                sim.addAgent((px[i], py[i]), *params, self.hrad[i] + 0.01 + c.orca_safety_space,
                             self.hvpref[i], (vx[i], vy[i]))
                for o in others:
                    sim.addAgent((o[0], o[1]), *params, o[4] + 0.01 + c.orca_safety_space, 1, (o[2], o[3]))
                self.sims[i] = sim
            else:
                sim.setAgentPosition(0, (px[i], py[i]))
                sim.setAgentVelocity(0, (vx[i], vy[i]))
                for k, o in enumerate(others):
                    sim.setAgentPosition(k + 1, (o[0], o[1]))
                    sim.setAgentVelocity(k + 1, (o[2], o[3]))
            velocity = np.array((self.hgx[i] - px[i], self.hgy[i] - py[i]))
            speed = np.linalg.norm(velocity)
            pref_vel = velocity / speed if speed > 1 else velocity
            sim.setAgentPrefVelocity(0, tuple(pref_vel))
            for k in range(len(others)):
                sim.setAgentPrefVelocity(k + 1, (0, 0))
            sim.doStep()
            acts.append(sim.getAgentVelocity(0))
            self.last_orca_diag.append((sim._numLines(0), sim._lineFail(0)))
        return acts

    def _social_force_actions(self, px, py, vx, vy, use_fov):
        """SOCIAL_FORCE.predict (crowd_nav/policy/social_force.py:11-49) for every human: pull towards the goal,
        exponential push from every other human (out-of-FOV ones replaced by the dummy at (7, 7)), speed clipped."""
        c, H = self.cfg, self.H
        fov = np.pi * c.human_fov
        acts = []
        for i in range(H):
            dx, dy = self.hgx[i] - px[i], self.hgy[i] - py[i]
            dist = np.sqrt(dx ** 2 + dy ** 2)
            dvx = c.sf_KI * ((dx / dist) * self.hvpref[i] - vx[i])
            dvy = c.sf_KI * ((dy / dist) * self.hvpref[i] - vy[i])
            ivx = ivy = 0
            for j in range(H):
                if j == i:
                    continue
                if (not use_fov) or self._in_fov(px[i], py[i], vx[i], vy[i], px[j], py[j], fov):
                    ox, oy, orad = px[j], py[j], self.hrad[j]
                else:
                    ox, oy, orad = 7, 7, 0.3
                ex, ey = px[i] - ox, py[i] - oy
                d = np.sqrt(ex ** 2 + ey ** 2)
                ivx += c.sf_A * np.exp((self.hrad[i] + orad - d) / c.sf_B) * (ex / d)
                ivy += c.sf_A * np.exp((self.hrad[i] + orad - d) / c.sf_B) * (ey / d)
            nvx = vx[i] + (dvx + ivx) * c.time_step
            nvy = vy[i] + (dvy + ivy) * c.time_step
            nrm = np.linalg.norm([nvx, nvy])
            if nrm > self.hvpref[i]:
                nvx, nvy = nvx / nrm * self.hvpref[i], nvy / nrm * self.hvpref[i]
            acts.append((nvx, nvy))
            self.last_orca_diag.append((0, -1))
        return acts

    def _truth_future_traj(self):
        """calc_human_future_traj('truth') (crowd_sim_var_num.py:152-228): buffer_len nested ORCA steps of all
        humans (the invisible robot does not take part), every pred_interval-th state kept, humans the robot
        does not currently see parked at (15, 15)."""
        c, H, P = self.cfg, self.H, self.P
        buffer_len = P * self.pred_interval
        traj = np.zeros((buffer_len + 1, H, 4))
        for i in range(H):
            traj[0, i] = [self.hpx[i], self.hpy[i], self.hvx[i], self.hvy[i]]
        for t in range(1, buffer_len + 1):
            acts = self._orca_actions(traj[t - 1, :, 0], traj[t - 1, :, 1], traj[t - 1, :, 2], traj[t - 1, :, 3],
                                      use_fov=False)
            for j, (ax, ay) in enumerate(acts):
                traj[t, j] = [traj[t - 1, j, 0] + ax * c.time_step, traj[t - 1, j, 1] + ay * c.time_step, ax, ay]
            self.last_sim_actions = acts
        traj = traj[::self.pred_interval]
        inv = np.logical_not(np.array(self.human_visibility, dtype=bool))
        traj[:, inv, :2] = 15
        traj[:, inv, 2:] = 0
        self.human_future_traj = traj
        return traj

    # ---------------------------------------------------------------- reward
    def _calc_reward(self):
        c, H = self.cfg, self.H
        dmin = float("inf")
        collision = False
        for i in range(H):
            dx = self.hpx[i] - self.rpx
            dy = self.hpy[i] - self.rpy
            closest = (dx ** 2 + dy ** 2) ** (1 / 2) - self.hrad[i] - c.robot_radius
            if closest < 0:
                collision = True
                break
            elif closest < dmin:
                dmin = closest
        reaching_goal = np.linalg.norm(np.array([self.rpx, self.rpy]) - np.array([self.rgx, self.rgy])) < c.robot_radius
        min_danger = 0.0
        if self.phase == "train":
            danger = dmin < c.discomfort_dist
        else:
            rel = self.human_future_traj[1:, :, :2] - np.array([self.rpx, self.rpy])
            rd = np.linalg.norm(rel, axis=-1)
            idx = rd < c.robot_radius + c.human_radius
            danger = bool(np.any(idx))
            if danger:
                min_danger = float(np.amin(rd[idx]))
        if self.global_time >= c.time_limit - 1:
            reward, done, info = 0, True, INFO_TIMEOUT
        elif collision:
            reward, done, info = c.collision_penalty, True, INFO_COLLISION
        elif reaching_goal:
            reward, done, info = c.success_reward, True, INFO_REACHGOAL
        elif danger:
            reward = (dmin - c.discomfort_dist) * c.discomfort_penalty_factor * c.time_step
            done, info = False, INFO_DANGER
        else:
            pot = np.linalg.norm(np.array([self.rpx, self.rpy]) - np.array([self.rgx, self.rgy]))
            reward = 2 * (-abs(pot) - self.potential)
            self.potential = -abs(pot)
            done, info = False, INFO_NOTHING
        if c.predict_method == "const_vel":
            # CrowdSimPred.calc_reward (crowd_sim_pred.py:216-233): STORED future trajectory
            rel = self.human_future_traj[1:, :, :2] - np.array([self.rpx, self.rpy])
            idx = np.linalg.norm(rel, axis=-1) < c.robot_radius + c.human_radius
            coef = 2. ** np.arange(2, self.P + 2).reshape((self.P, 1))
            reward = reward + np.min(idx * (c.collision_penalty / coef))
        if info != INFO_DANGER:
            min_danger = 0.0        # only Danger(min_dist) carries it (crowd_sim_var_num.py:526-532)
        return reward, done, info, min_danger

    # ---------------------------------------------------------------- step
    def step(self, action):
        """action: float32 array [2] (mutated in place like SRNN.clip_action, srnn.py:17-33)."""
        c, H = self.cfg, self.H
        act_norm = np.linalg.norm(action)
        if act_norm > c.robot_v_pref:
            action[0] = action[0] / act_norm * c.robot_v_pref
            action[1] = action[1] / act_norm * c.robot_v_pref
        avx, avy = action[0], action[1]
        human_actions = self._human_actions()
        # what reading the per-human simulators after the step shows (tools/make_golden.py does that): the LAST solve
        self.last_sim_actions = human_actions
        if self.phase == "test":
            self._truth_future_traj()          # crowd_sim_pred.py:136-138 / crowd_sim_var_num.py:386-388
        reward, done, info, min_danger = self._calc_reward()
        # integrate (agent.py:143-183)
        self.rpx = self.rpx + avx * c.time_step
        self.rpy = self.rpy + avy * c.time_step
        self.rvx, self.rvy = avx, avy
        for i, (vx, vy) in enumerate(human_actions):
            self.hpx[i] = self.hpx[i] + vx * c.time_step
            self.hpy[i] = self.hpy[i] + vy * c.time_step
            self.hvx[i], self.hvy[i] = vx, vy
        self.global_time += c.time_step
        self.step_counter += 1
        if c.human_num_range > 0 and self.global_time % 5 == 0:
            self._add_or_remove_humans()
            H = self.H
        ob = self._generate_ob(reset=False)
        if c.random_goal_changing and self.global_time % 5 == 0:
            self._update_goals_randomly()
        if c.end_goal_changing:
            for i in range(H):
                if _norm2(self.hgx[i] - self.hpx[i], self.hgy[i] - self.hpy[i]) < self.hrad[i]:
                    px, py, v_pref, radius = self._circle_crossing_human()
                    self.hpx[i], self.hpy[i], self.hgx[i], self.hgy[i] = px, py, -px, -py
                    self.hvx[i], self.hvy[i] = 0, 0
                    self.hvpref[i], self.hrad[i] = v_pref, radius
                    self.sims[i] = None     # new Human object => new ORCA policy => new rvo2 sim
        self.last_human_actions = human_actions
        return ob, reward, done, {"info": info, "min_danger": min_danger}

    def _add_or_remove_humans(self):
        """crowd_sim_var_num.py:404-437 (CrowdSimVarNum-v0) / crowd_sim_pred.py:165-194 (CrowdSimPred-v0): every 5 s
        at most human_num_range humans leave (the LAST ones, never below a currently observed id) or join."""
        c = self.cfg
        if self.rng.rand() < 0.5:
            seen = self.observed_human_ids          # CrowdSimPred never updates it after reset: always empty there
            if c.predict_method == "const_vel":
                max_remove = self.H - 1 if len(seen) == 0 else (self.H - 1) - max(seen)
                remove_num = int(self.rng.randint(low=0, high=min(c.human_num_range, max_remove) + 1))
            else:
                if len(seen) == 0:
                    max_remove = self.H - self.Hmin
                else:
                    max_remove = min(self.H - self.Hmin, (self.H - 1) - max(seen))
                remove_num = int(self.rng.randint(low=0, high=max_remove + 1))
            for _ in range(remove_num):
                for lst in (self.hpx, self.hpy, self.hvx, self.hvy, self.hgx, self.hgy, self.hrad, self.hvpref, self.sims):
                    lst.pop()
            self.H -= remove_num
            self.last_human_states = self.last_human_states[:self.H]
        else:
            add_num = int(self.rng.randint(low=0, high=c.human_num_range + 1))
            true_add = 0
            for i in range(self.H, self.H + add_num):
                if i == self.Hmax:
                    break
                px, py, v_pref, radius = self._circle_crossing_human()
                self.hpx.append(px); self.hpy.append(py); self.hvx.append(0); self.hvy.append(0)
                self.hgx.append(-px); self.hgy.append(-py); self.hrad.append(radius); self.hvpref.append(v_pref)
                self.sims.append(None)
                true_add += 1
            self.H += true_add
            if true_add > 0:
                self.last_human_states = np.concatenate((self.last_human_states, np.array([[15, 15, 0, 0, 0.3]] * true_add)), axis=0)
        assert 1 <= self.H <= self.Hmax

    def _update_goals_randomly(self):
        """crowd_sim.py:415-450."""
        c, H = self.cfg, self.H
        for i in range(H):
            if self.hvpref[i] == 0:
                continue
            if self.rng.random_sample() <= c.goal_change_chance:
                while True:
                    angle = self.rng.random_sample() * np.pi * 2
                    v_pref = 1.0 if self.hvpref[i] == 0 else self.hvpref[i]
                    gx_noise = (self.rng.random_sample() - 0.5) * v_pref
                    gy_noise = (self.rng.random_sample() - 0.5) * v_pref
                    gx = c.circle_radius * np.cos(angle) + gx_noise
                    gy = c.circle_radius * np.sin(angle) + gy_noise
                    collide = False
                    for k in range(-1, H):
                        if k == i:
                            continue
                        if k < 0:
                            ax, ay, agx, agy, ar = self.rpx, self.rpy, self.rgx, self.rgy, c.robot_radius
                        else:
                            ax, ay, agx, agy, ar = self.hpx[k], self.hpy[k], self.hgx[k], self.hgy[k], self.hrad[k]
                        min_dist = self.hrad[i] + ar + c.discomfort_dist
                        if _norm2(gx - ax, gy - ay) < min_dist or _norm2(gx - agx, gy - agy) < min_dist:
                            collide = True
                            break
                    if not collide:
                        break
                self.hgx[i], self.hgy[i] = gx, gy

    # ---------------------------------------------------------------- worker semantics
    def worker_step(self, action):
        """rl/networks/shmem_vec_env.py:138-142 + bench.Monitor: step, auto-reset on done."""
        ob, reward, done, info = self.step(action)
        self.ep_ret += float(reward)
        self.ep_len += 1
        if done:
            info["episode"] = {"r": round(self.ep_ret, 6), "l": self.ep_len}
            ob = self.reset()
        return ob, reward, done, info

    # ---------------------------------------------------------------- state export (parity tests)
    def get_state(self):
        H = self.H
        f = lambda x: np.asarray(x, dtype=np.float64)
        return {
            "robot": f([self.rpx, self.rpy, self.rvx, self.rvy, self.rgx, self.rgy]),
            "hpx": f(self.hpx), "hpy": f(self.hpy), "hvx": f(self.hvx), "hvy": f(self.hvy),
            "hgx": f(self.hgx), "hgy": f(self.hgy), "hrad": f(self.hrad), "hvpref": f(self.hvpref),
            "belief": self.last_human_states.copy(),
            "traj": None if self.human_future_traj is None else self.human_future_traj.copy(),
            "vis": np.asarray(self.human_visibility, dtype=bool),
            "global_time": float(self.global_time), "potential": float(self.potential),
            "nd_global": float(self.nd_global),
            "case_counter": int(self.case_counter[self.phase]),
            "sim_exists": np.asarray([s is not None for s in self.sims], dtype=bool),
        }


class OracleVecEnv(object):
    """N oracle environments stepped serially with the VecEnv contract of
    rl/networks/shmem_vec_env.py (obs float32 [N,...], rewards float64 [N], dones bool [N])."""

    def __init__(self, cfg, num_envs, seed=425, rank_offset=0, nenv_total=None, phase="train"):
        total = num_envs if nenv_total is None else nenv_total
        self.envs = [CrowdEnvOracle(cfg, seed + rank_offset + k, total, phase) for k in range(num_envs)]
        self.num_envs = num_envs

    @staticmethod
    def _stack(obs):
        return {k: np.stack([o[k] for o in obs]) for k in obs[0]}

    def reset(self):
        return self._stack([e.reset() for e in self.envs])

    def step(self, actions):
        actions = np.array(actions, dtype=np.float32, copy=True)
        outs = [e.worker_step(actions[k]) for k, e in enumerate(self.envs)]
        obs, rews, dones, infos = zip(*outs)
        return self._stack(obs), np.array(rews, dtype=np.float64), np.array(dones), infos
