// selftest.cpp -- pins the oracle (TEST INFRASTRUCTURE) against the reference's own eight
// tests, restated with the reference's constants, plus the config-C1 scene of
// rmf_crowdsim_viz/src/main.rs:64-94 as a smoke run.  Exit code 0 = all pass.
//
//   lib.rs:422-453                         test_step_integration
//   location_hash_2d.rs:310-339            test_nearest_neighbours
//   location_hash_2d.rs:342-368            test_radius_search
//   location_hash_2d.rs:370-381            test_update
//   location_hash_2d.rs:384-397            test_remove
//   zanlungo.rs:224-229                    test_time_to_collision_head_on
//   zanlungo.rs:231-236                    test_time_to_collision_never_collide
//   tests/event_listeners_test.rs:64-111   test_event_listener_source_sink_api
#include <cstdio>
#include <cstdlib>
#include <set>

#include "crowdsim_oracle.hpp"

using namespace orc;

static int g_fail = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);          \
      g_fail++;                                                            \
    }                                                                      \
  } while (0)

static void fill_100(LocationHash2D& h, std::unordered_map<AgentId, Vec2>& naive) {
  AgentId id = 0;
  for (int x = 0; x < 10; ++x)
    for (int y = 0; y < 10; ++y) {
      Vec2 p{x + 0.5, y + 0.5};
      CHECK(h.add_or_update(id, p).ok);
      naive[id] = p;
      id++;
    }
}

static void test_nearest_neighbours() {
  LocationHash2D h(10, 10, 0.5, {0, 0});
  std::unordered_map<AgentId, Vec2> naive;
  fill_100(h, naive);
  auto nb = h.get_nearest_neighbours(1, {0.6, 0.6});
  CHECK(nb.size() == 1 && nb[0] == 0);
  nb = h.get_nearest_neighbours(4, {1.7, 1.6});
  // brute force (location_hash_2d.rs:275-292)
  std::vector<std::pair<double, AgentId>> all;
  for (auto& kv : naive) all.push_back({norm(kv.second - Vec2{1.7, 1.6}), kv.first});
  std::stable_sort(all.begin(), all.end(), [](auto& a, auto& b) { return a.first < b.first; });
  std::vector<AgentId> gt;
  for (int i = 0; i < 4; ++i) gt.push_back(all[i].second);
  CHECK(nb == gt);
  CHECK((nb == std::vector<AgentId>{11, 21, 12, 10}));
}

static void test_radius_search() {
  LocationHash2D h(10, 10, 0.5, {0, 0});
  std::unordered_map<AgentId, Vec2> naive;
  fill_100(h, naive);
  std::set<AgentId> gt;
  for (auto& kv : naive)
    if (norm(kv.second - Vec2{4, 4}) < 1.1) gt.insert(kv.first);
  auto nb = h.get_neighbours_in_radius(1.1, {4, 4});
  std::set<AgentId> got(nb.begin(), nb.end());
  CHECK(got == gt);
  CHECK((got == std::set<AgentId>{33, 34, 43, 44}));
}

static void test_update() {
  LocationHash2D h(2, 2, 1, {0, 0});
  h.add_or_update(1, {0, 0});
  auto a = h.get_neighbours_in_radius(1, {0, 0});
  CHECK(a.size() == 1 && a[0] == 1);
  h.add_or_update(1, {1, 0});
  a = h.get_neighbours_in_radius(1, {0, 0});
  CHECK(a.size() == 0);
}

static void test_remove() {
  LocationHash2D h(1, 1, 1, {0, 0});
  h.add_or_update(1, {0, 0});
  CHECK(h.get_neighbours_in_radius(1.1, {0, 0}).size() == 1);
  h.remove_agent(1);
  CHECK(h.get_neighbours_in_radius(1.1, {0, 0}).size() == 0);
}

static void test_ttc() {
  Zanlungo z(1, 10, 0, 5, 0.1, 4);
  CHECK(z.time_to_collision({1, 0}, {-10, 0}) == 6.0);
  CHECK(z.time_to_collision({1, 0}, {10, 0}) == std::numeric_limits<double>::infinity());
}

static void test_step_integration() {
  for (int mode = 0; mode < 2; ++mode) {
    Simulation sim(std::make_unique<LocationHash2D>(1000, 1000, 20, Vec2{-500, -500}),
                   static_cast<IndexMode>(mode));
    Locked<HighLevelPlanner> hl(std::make_shared<ConstantVelocityPlan>(Vec2{1, 0}));
    Locked<LocalPlanner> lp(std::make_shared<NoLocalPlan>());
    CHECK(sim.agents.size() == 0);
    std::vector<AgentId> ids;
    CHECK(sim.add_agents({Vec2{0, 0}}, hl, lp, 100, &ids).ok);
    CHECK(ids.size() == 1 && sim.agents.size() == 1);
    CHECK(sim.step(Duration{1, 0}).ok);
    CHECK(sim.agents.size() == 1);
    CHECK(norm(sim.agents.at(0).position - Vec2{1, 0}) < 1e-5);
  }
}

struct MockEventListener : EventListener {
  std::vector<AgentId> added, removed;
  void agent_spawned(Vec2, AgentId a) override { added.push_back(a); }
  void agent_destroyed(AgentId a) override { removed.push_back(a); }
};

static void test_event_listener_source_sink_api() {
  for (int mode = 0; mode < 2; ++mode) {
    Simulation sim(std::make_unique<LocationHash2D>(1000, 1000, 20, Vec2{-500, -500}),
                   static_cast<IndexMode>(mode));
    auto ss = std::make_shared<SourceSink>();
    ss->source = {0, 0};
    ss->waypoints = {Vec2{20, 0}};
    ss->radius_sink = 1;
    ss->crowd_generator = std::make_shared<MonotonicCrowd>(1.0);
    ss->high_level_planner = Locked<HighLevelPlanner>(std::make_shared<ConstantVelocityPlan>(Vec2{1, 0}));
    ss->local_planner = Locked<LocalPlanner>(std::make_shared<NoLocalPlan>());
    ss->agent_eyesight_range = 5;
    ss->loop_forever = false;
    auto listener = std::make_shared<MockEventListener>();
    sim.add_event_listener(listener);
    sim.add_source_sink(ss);
    for (size_t steps = 0; steps < 20; ++steps) {
      CHECK(sim.agents.size() == steps);
      CHECK(listener->added.size() == steps);
      sim.step(Duration{1, 0});
    }
    for (size_t steps = 20; steps < 40; ++steps) {
      CHECK(sim.agents.size() == 20);
      CHECK(listener->added.size() == steps);
      CHECK(listener->removed.size() == steps - 20);
      sim.step(Duration{1, 0});
    }
  }
}

// Config C1: the viz scene (main.rs:64-94) at dt = 16_666_667 ns.  The expected values are
// SURVEY.md section 8(c)'s survey-derived cross-check (a throw-away Python restatement), NOT a
// reference test; they are compared loosely (1e-9 relative) and only as a smoke check.
static bool close(double a, double b, double rel = 1e-9) {
  return std::fabs(a - b) <= rel * std::max(1.0, std::max(std::fabs(a), std::fabs(b)));
}
static void test_c1_scene() {
  Simulation sim(std::make_unique<LocationHash2D>(1000, 1000, 20, Vec2{-500, -500}));
  Locked<HighLevelPlanner> hl(std::make_shared<ParityVelocityPlan>(Vec2{0, 10}));
  auto z = std::make_shared<Zanlungo>(1, 1, 0, 40, 2, 20);
  Locked<LocalPlanner> lp(z);
  CHECK(sim.add_agents({Vec2{100, 100}, Vec2{100, -100}, Vec2{60, 100}}, hl, lp, 100, nullptr).ok);
  sim.enable_trace(true);
  Duration dt{0, 16666667};
  for (int s = 1; s <= 1000; ++s) {
    CHECK(sim.step(dt).ok);
    if (s == 1) {
      CHECK(close(sim.agents.at(0).position.y, 99.83333333));
      for (auto& t : sim.trace()) CHECK(std::isinf(t.t_i));
    }
    if (s == 301) {
      const AgentTrace* t0 = nullptr;
      const AgentTrace* t1 = nullptr;
      for (auto& t : sim.trace()) {
        if (t.id == 0) t0 = &t;
        if (t.id == 1) t1 = &t;
      }
      CHECK(t0 && t1);
      CHECK(close(t0->t_i, 3.999999900000064));
      CHECK(close(t0->force.x, -3.0326534501957414) && close(t0->force.y, -3.304299148053835));
      CHECK(close(sim.agents.at(0).velocity.x, -1.5163267250978707));
      CHECK(close(sim.agents.at(0).velocity.y, -11.652149574026918));
      CHECK(close(sim.agents.at(0).position.x, 99.9747278874096));
      CHECK(close(sim.agents.at(0).position.y, 49.80579650321613));
      CHECK(close(t1->t_i, 3.999999900000064));
      CHECK(t1->force.x == 0.0 && t1->force.y == 0.0);
      CHECK(sim.agents.at(1).velocity.x == 0.0 && sim.agents.at(1).velocity.y == 10.0);
    }
    if (s == 302) {
      for (auto& t : sim.trace())
        if (t.id == 0) {
          CHECK(close(t.t_i, 3.7158738861356664));
          CHECK(close(t.force.x, -4.176664976953017) && close(t.force.y, -4.316993668385091));
        }
    }
  }
  CHECK(close(sim.agents.at(0).position.x, 84.67547281746326, 1e-7));
  CHECK(close(sim.agents.at(0).position.y, -82.20914817564514, 1e-7));
  CHECK(close(sim.agents.at(1).position.x, 104.94348286863043, 1e-7));
  CHECK(close(sim.agents.at(1).position.y, 70.9351303745993, 1e-7));
  CHECK(close(sim.agents.at(2).position.x, 60.0, 1e-7));
  CHECK(close(sim.agents.at(2).position.y, -66.66666999999896, 1e-7));
  std::printf("C1 step 1000: p0=(%.17g, %.17g) p1=(%.17g, %.17g) p2=(%.17g, %.17g)\n",
              sim.agents.at(0).position.x, sim.agents.at(0).position.y, sim.agents.at(1).position.x,
              sim.agents.at(1).position.y, sim.agents.at(2).position.x, sim.agents.at(2).position.y);
}

int main() {
  test_nearest_neighbours();
  test_radius_search();
  test_update();
  test_remove();
  test_ttc();
  test_step_integration();
  test_event_listener_source_sink_api();
  test_c1_scene();
  if (g_fail == 0) std::printf("oracle selftest: ALL PASS\n");
  return g_fail == 0 ? 0 : 1;
}
