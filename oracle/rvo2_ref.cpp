// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the RVO2 library (ORCA, van den Berg / Guy / Snape / Lin / Manocha,
// "Reciprocal n-body collision avoidance", RVO2 Library v2.0.x) as it is driven by the
// reference through the un-vendored Cython wrapper `rvo2` (sybrenstuvel/Python-RVO2).
//
// PARITY UNPINNED: the reference names this dependency only in README.md:39, pins no
// version and vendors no source, and has no tests or golden vectors at this boundary
// (SURVEY.md §8c).  This file restates the published algorithm (all arithmetic in
// `float`, RVO_EPSILON = 1e-5f, kd-tree leaf size 10) and is anchored on the reference's
// call sites:
//   crowd_nav/policy/orca.py:84      PyRVOSimulator(timeStep, neighborDist, maxNeighbors,
//                                                   timeHorizon, timeHorizonObst, radius, maxSpeed)
//   crowd_nav/policy/orca.py:85-89   addAgent(pos, neighborDist, maxNeighbors, timeHorizon,
//                                             timeHorizonObst, radius, maxSpeed, velocity)
//   crowd_nav/policy/orca.py:91-95   setAgentPosition / setAgentVelocity
//   crowd_nav/policy/orca.py:108-111 setAgentPrefVelocity
//   crowd_nav/policy/orca.py:113-114 doStep / getAgentVelocity
//   crowd_nav/policy/orca.py:80      getNumAgents
// No obstacle is ever added by the reference, so the obstacle half of RVO2 is omitted.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: no FMA contraction, so that the
// float sequence is the plain IEEE one the CUDA kernel reproduces with __fmul_rn/__fadd_rn).
//
// The C ABI below is bound by oracle/shims/rvo2.py (ctypes) as class PyRVOSimulator.

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>
#include <algorithm>

namespace {

constexpr float kEps = 0.00001f;       // RVO_EPSILON
constexpr size_t kMaxLeaf = 10;        // KdTree MAX_LEAF_SIZE

struct V2 {
  float x, y;
};
inline V2 mk(float x, float y) { return V2{x, y}; }
inline V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
inline V2 operator-(V2 a) { return mk(-a.x, -a.y); }
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
inline V2 operator*(V2 a, float s) { return mk(a.x * s, a.y * s); }
// RVO2's Vector2::operator/ multiplies by the reciprocal.
inline V2 operator/(V2 a, float s) {
  const float inv = 1.0f / s;
  return mk(a.x * inv, a.y * inv);
}
inline float absSq(V2 a) { return dot(a, a); }
inline float vabs(V2 a) { return std::sqrt(dot(a, a)); }
inline float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
inline V2 normalize(V2 a) { return a / vabs(a); }
inline float sqr(float a) { return a * a; }

struct Line {
  V2 point, direction;
};

struct Agent {
  V2 position{0, 0}, velocity{0, 0}, prefVelocity{0, 0}, newVelocity{0, 0};
  float neighborDist = 0, timeHorizon = 0, timeHorizonObst = 0, radius = 0, maxSpeed = 0;
  size_t maxNeighbors = 0;
  size_t id = 0;
  std::vector<std::pair<float, const Agent*>> nbrs;
  std::vector<Line> lines;
  int lastLineFail = -1;   // diagnostics: -1 = LP2 succeeded
};

bool lp1(const std::vector<Line>& lines, size_t lineNo, float radius, V2 optVelocity,
         bool directionOpt, V2& result) {
  const float dotProduct = dot(lines[lineNo].point, lines[lineNo].direction);
  const float discriminant = sqr(dotProduct) + sqr(radius) - absSq(lines[lineNo].point);
  if (discriminant < 0.0f) return false;
  const float sqrtDiscriminant = std::sqrt(discriminant);
  float tLeft = -dotProduct - sqrtDiscriminant;
  float tRight = -dotProduct + sqrtDiscriminant;
  for (size_t i = 0; i < lineNo; ++i) {
    const float denominator = det(lines[lineNo].direction, lines[i].direction);
    const float numerator = det(lines[i].direction, lines[lineNo].point - lines[i].point);
    if (std::fabs(denominator) <= kEps) {
      if (numerator < 0.0f) return false;
      continue;
    }
    const float t = numerator / denominator;
    if (denominator >= 0.0f) {
      tRight = std::min(tRight, t);
    } else {
      tLeft = std::max(tLeft, t);
    }
    if (tLeft > tRight) return false;
  }
  if (directionOpt) {
    if (dot(optVelocity, lines[lineNo].direction) > 0.0f) {
      result = lines[lineNo].point + tRight * lines[lineNo].direction;
    } else {
      result = lines[lineNo].point + tLeft * lines[lineNo].direction;
    }
  } else {
    const float t = dot(lines[lineNo].direction, optVelocity - lines[lineNo].point);
    if (t < tLeft) {
      result = lines[lineNo].point + tLeft * lines[lineNo].direction;
    } else if (t > tRight) {
      result = lines[lineNo].point + tRight * lines[lineNo].direction;
    } else {
      result = lines[lineNo].point + t * lines[lineNo].direction;
    }
  }
  return true;
}

size_t lp2(const std::vector<Line>& lines, float radius, V2 optVelocity, bool directionOpt,
           V2& result) {
  if (directionOpt) {
    result = optVelocity * radius;
  } else if (absSq(optVelocity) > sqr(radius)) {
    result = normalize(optVelocity) * radius;
  } else {
    result = optVelocity;
  }
  for (size_t i = 0; i < lines.size(); ++i) {
    if (det(lines[i].direction, lines[i].point - result) > 0.0f) {
      const V2 tempResult = result;
      if (!lp1(lines, i, radius, optVelocity, directionOpt, result)) {
        result = tempResult;
        return i;
      }
    }
  }
  return lines.size();
}

void lp3(const std::vector<Line>& lines, size_t numObstLines, size_t beginLine, float radius,
         V2& result) {
  float distance = 0.0f;
  for (size_t i = beginLine; i < lines.size(); ++i) {
    if (det(lines[i].direction, lines[i].point - result) > distance) {
      std::vector<Line> projLines(lines.begin(),
                                  lines.begin() + static_cast<std::ptrdiff_t>(numObstLines));
      for (size_t j = numObstLines; j < i; ++j) {
        Line line;
        const float determinant = det(lines[i].direction, lines[j].direction);
        if (std::fabs(determinant) <= kEps) {
          if (dot(lines[i].direction, lines[j].direction) > 0.0f) {
            continue;
          }
          line.point = 0.5f * (lines[i].point + lines[j].point);
        } else {
          line.point = lines[i].point +
                       (det(lines[j].direction, lines[i].point - lines[j].point) / determinant) *
                           lines[i].direction;
        }
        line.direction = normalize(lines[j].direction - lines[i].direction);
        projLines.push_back(line);
      }
      const V2 tempResult = result;
      if (lp2(projLines, radius, mk(-lines[i].direction.y, lines[i].direction.x), true, result) <
          projLines.size()) {
        result = tempResult;
      }
      distance = det(lines[i].direction, lines[i].point - result);
    }
  }
}

struct Sim {
  float timeStep = 0;
  // defaults for addAgent without explicit parameters (unused by the reference)
  float defNeighborDist = 0, defTimeHorizon = 0, defTimeHorizonObst = 0, defRadius = 0,
        defMaxSpeed = 0;
  size_t defMaxNeighbors = 0;
  float globalTime = 0;
  std::vector<Agent*> agents;

  // kd-tree (agents only)
  struct Node {
    size_t begin, end, left, right;
    float maxX, maxY, minX, minY;
  };
  std::vector<Agent*> kdAgents;   // persists across steps, like KdTree::agents_
  std::vector<Node> tree;

  ~Sim() {
    for (Agent* a : agents) delete a;
  }

  void buildTree() {
    if (kdAgents.size() < agents.size()) {
      for (size_t i = kdAgents.size(); i < agents.size(); ++i) kdAgents.push_back(agents[i]);
      tree.resize(2 * kdAgents.size() - 1);
    }
    if (!kdAgents.empty()) buildRec(0, kdAgents.size(), 0);
  }

  void buildRec(size_t begin, size_t end, size_t node) {
    Node& n = tree[node];
    n.begin = begin;
    n.end = end;
    n.minX = n.maxX = kdAgents[begin]->position.x;
    n.minY = n.maxY = kdAgents[begin]->position.y;
    for (size_t i = begin + 1; i < end; ++i) {
      n.maxX = std::max(n.maxX, kdAgents[i]->position.x);
      n.minX = std::min(n.minX, kdAgents[i]->position.x);
      n.maxY = std::max(n.maxY, kdAgents[i]->position.y);
      n.minY = std::min(n.minY, kdAgents[i]->position.y);
    }
    if (end - begin > kMaxLeaf) {
      const bool isVertical = (n.maxX - n.minX > n.maxY - n.minY);
      const float splitValue =
          (isVertical ? 0.5f * (n.maxX + n.minX) : 0.5f * (n.maxY + n.minY));
      size_t left = begin, right = end;
      while (left < right) {
        while (left < right &&
               (isVertical ? kdAgents[left]->position.x : kdAgents[left]->position.y) < splitValue)
          ++left;
        while (right > left && (isVertical ? kdAgents[right - 1]->position.x
                                           : kdAgents[right - 1]->position.y) >= splitValue)
          --right;
        if (left < right) {
          std::swap(kdAgents[left], kdAgents[right - 1]);
          ++left;
          --right;
        }
      }
      if (left == begin) {
        ++left;
        ++right;
      }
      tree[node].left = node + 1;
      tree[node].right = node + 2 * (left - begin);
      const size_t l = tree[node].left, r = tree[node].right;
      buildRec(begin, left, l);
      buildRec(left, end, r);
    }
  }

  static void insertNeighbor(Agent* self, const Agent* other, float& rangeSq) {
    if (self == other) return;
    const float distSq = absSq(self->position - other->position);
    if (distSq < rangeSq) {
      if (self->nbrs.size() < self->maxNeighbors) self->nbrs.push_back({distSq, other});
      size_t i = self->nbrs.size() - 1;
      while (i != 0 && distSq < self->nbrs[i - 1].first) {
        self->nbrs[i] = self->nbrs[i - 1];
        --i;
      }
      self->nbrs[i] = {distSq, other};
      if (self->nbrs.size() == self->maxNeighbors) rangeSq = self->nbrs.back().first;
    }
  }

  float boxDistSq(const Node& c, const Agent* a) const {
    return sqr(std::max(0.0f, c.minX - a->position.x)) + sqr(std::max(0.0f, a->position.x - c.maxX)) +
           sqr(std::max(0.0f, c.minY - a->position.y)) + sqr(std::max(0.0f, a->position.y - c.maxY));
  }

  void queryRec(Agent* a, float& rangeSq, size_t node) const {
    const Node& n = tree[node];
    if (n.end - n.begin <= kMaxLeaf) {
      for (size_t i = n.begin; i < n.end; ++i) insertNeighbor(a, kdAgents[i], rangeSq);
      return;
    }
    const float dl = boxDistSq(tree[n.left], a);
    const float dr = boxDistSq(tree[n.right], a);
    if (dl < dr) {
      if (dl < rangeSq) {
        queryRec(a, rangeSq, n.left);
        if (dr < rangeSq) queryRec(a, rangeSq, n.right);
      }
    } else {
      if (dr < rangeSq) {
        queryRec(a, rangeSq, n.right);
        if (dl < rangeSq) queryRec(a, rangeSq, n.left);
      }
    }
  }

  void computeNeighbors(Agent* a) const {
    a->nbrs.clear();
    if (a->maxNeighbors > 0) {
      float rangeSq = sqr(a->neighborDist);
      queryRec(a, rangeSq, 0);
    }
  }

  void computeNewVelocity(Agent* a) const {
    a->lines.clear();
    const size_t numObstLines = 0;
    const float invTimeHorizon = 1.0f / a->timeHorizon;
    for (size_t i = 0; i < a->nbrs.size(); ++i) {
      const Agent* other = a->nbrs[i].second;
      const V2 relativePosition = other->position - a->position;
      const V2 relativeVelocity = a->velocity - other->velocity;
      const float distSq = absSq(relativePosition);
      const float combinedRadius = a->radius + other->radius;
      const float combinedRadiusSq = sqr(combinedRadius);
      Line line;
      V2 u;
      if (distSq > combinedRadiusSq) {
        const V2 w = relativeVelocity - invTimeHorizon * relativePosition;
        const float wLengthSq = absSq(w);
        const float dotProduct1 = dot(w, relativePosition);
        if (dotProduct1 < 0.0f && sqr(dotProduct1) > combinedRadiusSq * wLengthSq) {
          const float wLength = std::sqrt(wLengthSq);
          const V2 unitW = w / wLength;
          line.direction = mk(unitW.y, -unitW.x);
          u = (combinedRadius * invTimeHorizon - wLength) * unitW;
        } else {
          const float leg = std::sqrt(distSq - combinedRadiusSq);
          if (det(relativePosition, w) > 0.0f) {
            line.direction = mk(relativePosition.x * leg - relativePosition.y * combinedRadius,
                                relativePosition.x * combinedRadius + relativePosition.y * leg) /
                             distSq;
          } else {
            line.direction = -mk(relativePosition.x * leg + relativePosition.y * combinedRadius,
                                 -relativePosition.x * combinedRadius + relativePosition.y * leg) /
                             distSq;
          }
          const float dotProduct2 = dot(relativeVelocity, line.direction);
          u = dotProduct2 * line.direction - relativeVelocity;
        }
      } else {
        const float invTimeStep = 1.0f / timeStep;
        const V2 w = relativeVelocity - invTimeStep * relativePosition;
        const float wLength = vabs(w);
        const V2 unitW = w / wLength;
        line.direction = mk(unitW.y, -unitW.x);
        u = (combinedRadius * invTimeStep - wLength) * unitW;
      }
      line.point = a->velocity + 0.5f * u;
      a->lines.push_back(line);
    }
    const size_t lineFail = lp2(a->lines, a->maxSpeed, a->prefVelocity, false, a->newVelocity);
    a->lastLineFail = -1;
    if (lineFail < a->lines.size()) {
      a->lastLineFail = static_cast<int>(lineFail);
      lp3(a->lines, numObstLines, lineFail, a->maxSpeed, a->newVelocity);
    }
  }

  // full RVOSimulator::doStep: every agent gets a new velocity and is integrated.
  void doStep() {
    buildTree();
    for (Agent* a : agents) {
      computeNeighbors(a);
      computeNewVelocity(a);
    }
    for (Agent* a : agents) {
      a->velocity = a->newVelocity;
      a->position = a->position + a->velocity * timeStep;
    }
    globalTime += timeStep;
  }

  // test-speed variant: only agent `idx` is solved (its result does not depend on the
  // other agents' solves because doStep reads positions/velocities before any update).
  void doStepOnly(size_t idx) {
    buildTree();
    Agent* a = agents[idx];
    computeNeighbors(a);
    computeNewVelocity(a);
    a->velocity = a->newVelocity;
    a->position = a->position + a->velocity * timeStep;
    globalTime += timeStep;
  }
};

}  // namespace

extern "C" {

void* rvo2ref_create(float timeStep, float neighborDist, int maxNeighbors, float timeHorizon,
                     float timeHorizonObst, float radius, float maxSpeed) {
  Sim* s = new Sim();
  s->timeStep = timeStep;
  s->defNeighborDist = neighborDist;
  s->defMaxNeighbors = static_cast<size_t>(maxNeighbors);
  s->defTimeHorizon = timeHorizon;
  s->defTimeHorizonObst = timeHorizonObst;
  s->defRadius = radius;
  s->defMaxSpeed = maxSpeed;
  return s;
}

void rvo2ref_destroy(void* h) { delete static_cast<Sim*>(h); }

int rvo2ref_add_agent(void* h, float px, float py, float neighborDist, int maxNeighbors,
                      float timeHorizon, float timeHorizonObst, float radius, float maxSpeed,
                      float vx, float vy) {
  Sim* s = static_cast<Sim*>(h);
  Agent* a = new Agent();
  a->position = mk(px, py);
  a->velocity = mk(vx, vy);
  a->neighborDist = neighborDist;
  a->maxNeighbors = static_cast<size_t>(maxNeighbors);
  a->timeHorizon = timeHorizon;
  a->timeHorizonObst = timeHorizonObst;
  a->radius = radius;
  a->maxSpeed = maxSpeed;
  a->id = s->agents.size();
  s->agents.push_back(a);
  return static_cast<int>(a->id);
}

int rvo2ref_num_agents(void* h) { return static_cast<int>(static_cast<Sim*>(h)->agents.size()); }

void rvo2ref_set_position(void* h, int i, float x, float y) {
  static_cast<Sim*>(h)->agents[static_cast<size_t>(i)]->position = mk(x, y);
}
void rvo2ref_set_velocity(void* h, int i, float x, float y) {
  static_cast<Sim*>(h)->agents[static_cast<size_t>(i)]->velocity = mk(x, y);
}
void rvo2ref_set_pref_velocity(void* h, int i, float x, float y) {
  static_cast<Sim*>(h)->agents[static_cast<size_t>(i)]->prefVelocity = mk(x, y);
}
void rvo2ref_get_velocity(void* h, int i, float* out) {
  const Agent* a = static_cast<Sim*>(h)->agents[static_cast<size_t>(i)];
  out[0] = a->velocity.x;
  out[1] = a->velocity.y;
}
void rvo2ref_get_position(void* h, int i, float* out) {
  const Agent* a = static_cast<Sim*>(h)->agents[static_cast<size_t>(i)];
  out[0] = a->position.x;
  out[1] = a->position.y;
}
void rvo2ref_do_step(void* h) { static_cast<Sim*>(h)->doStep(); }
void rvo2ref_do_step_only(void* h, int i) { static_cast<Sim*>(h)->doStepOnly(static_cast<size_t>(i)); }

// diagnostics for parity tests: number of ORCA lines built for agent i in the last
// doStep, the LP2 failure line (-1 when LP2 succeeded), and the neighbour order.
int rvo2ref_num_lines(void* h, int i) {
  return static_cast<int>(static_cast<Sim*>(h)->agents[static_cast<size_t>(i)]->lines.size());
}
int rvo2ref_line_fail(void* h, int i) {
  return static_cast<Sim*>(h)->agents[static_cast<size_t>(i)]->lastLineFail;
}
int rvo2ref_neighbor_ids(void* h, int i, int* out, int cap) {
  const Agent* a = static_cast<Sim*>(h)->agents[static_cast<size_t>(i)];
  int n = 0;
  for (const auto& p : a->nbrs) {
    if (n < cap) out[n] = static_cast<int>(p.second->id);
    ++n;
  }
  return n;
}
void rvo2ref_get_line(void* h, int i, int k, float* out4) {
  const Line& l = static_cast<Sim*>(h)->agents[static_cast<size_t>(i)]->lines[static_cast<size_t>(k)];
  out4[0] = l.point.x;
  out4[1] = l.point.y;
  out4[2] = l.direction.x;
  out4[3] = l.direction.y;
}

// Batched helper used by oracle/crowd_env.py for the bounded CPU baseline and by the
// parity tests: ONE ego agent against `n_other` others (ego = agent 0, exactly the
// per-human simulator the reference builds at crowd_nav/policy/orca.py:84-95), fresh
// simulator each call (so kd-tree order = insertion order).  Arrays are float32.
//   ego: px,py,vx,vy,radius,max_speed,pref_x,pref_y ; others: n_other x (px,py,vx,vy,radius)
// returns new velocity in out[0..1]; out_diag = {num_lines, line_fail}.
void rvo2ref_solve_one(const float* ego, const float* others, int n_other, float neighborDist,
                       float timeHorizon, float timeStep, float otherMaxSpeed, float* out,
                       int* out_diag) {
  Sim s;
  s.timeStep = timeStep;
  rvo2ref_add_agent(&s, ego[0], ego[1], neighborDist, n_other, timeHorizon, timeHorizon, ego[4],
                    ego[5], ego[2], ego[3]);
  for (int j = 0; j < n_other; ++j) {
    const float* o = others + 5 * j;
    rvo2ref_add_agent(&s, o[0], o[1], neighborDist, n_other, timeHorizon, timeHorizon, o[4],
                      otherMaxSpeed, o[2], o[3]);
  }
  s.agents[0]->prefVelocity = mk(ego[6], ego[7]);
  s.doStepOnly(0);
  out[0] = s.agents[0]->velocity.x;
  out[1] = s.agents[0]->velocity.y;
  if (out_diag) {
    out_diag[0] = static_cast<int>(s.agents[0]->lines.size());
    out_diag[1] = s.agents[0]->lastLineFail;
  }
}

}  // extern "C"
