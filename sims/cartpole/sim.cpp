#include "sim.hpp"

#ifdef MADRONA_GPU_MODE
#include <madrona/mw_gpu_entry.hpp>
#endif

using namespace madrona;

namespace cartpole {

void Sim::registerTypes(ECSRegistry &registry, const Config &)
{
    registry.registerComponent<Action>();
    registry.registerComponent<CartState>();
    registry.registerComponent<Reward>();
    registry.registerComponent<Done>();
    registry.registerComponent<StepCount>();
    registry.registerSingleton<WorldReset>();

    registry.registerArchetype<Cart>(
        ComponentMetadataSelector<> {}, ArchetypeFlags::None, 1);

    registry.exportSingleton<WorldReset>((uint32_t)ExportID::Reset);
    registry.exportColumn<Cart, Action>((uint32_t)ExportID::Action);
    registry.exportColumn<Cart, CartState>((uint32_t)ExportID::State);
    registry.exportColumn<Cart, Reward>((uint32_t)ExportID::Reward);
    registry.exportColumn<Cart, Done>((uint32_t)ExportID::Done);
}

static inline void resetCart(Engine &ctx, CartState &s, StepCount &count)
{
    Sim &sim = ctx.data();
    // uniform in [-0.05, 0.05)
    s.x = sim.rng.sampleUniform() * 0.1f - 0.05f;
    s.xDot = sim.rng.sampleUniform() * 0.1f - 0.05f;
    s.theta = sim.rng.sampleUniform() * 0.1f - 0.05f;
    s.thetaDot = sim.rng.sampleUniform() * 0.1f - 0.05f;
    count.t = 0;
    sim.episode += 1;
}

// Degree-7 / degree-6 Taylor polynomials: this fixture's own definition of
// the trigonometric terms (|theta| stays below 0.21 rad before termination).
static inline float polySin(float t)
{
    float t2 = t * t;
    return t * (1.f - t2 * (1.f / 6.f - t2 * (1.f / 120.f - t2 * (1.f / 5040.f))));
}

static inline float polyCos(float t)
{
    float t2 = t * t;
    return 1.f - t2 * (0.5f - t2 * (1.f / 24.f - t2 * (1.f / 720.f)));
}

inline void stepSystem(Engine &ctx,
                       Action &action,
                       CartState &s,
                       Reward &reward,
                       Done &done,
                       StepCount &count)
{
    WorldReset &reset = ctx.singleton<WorldReset>();
    if (reset.reset != 0 || done.v != 0) {
        resetCart(ctx, s, count);
        reset.reset = 0;
    }

    constexpr float gravity = 9.8f;
    constexpr float mass_cart = 1.0f;
    constexpr float mass_pole = 0.1f;
    constexpr float total_mass = mass_cart + mass_pole;
    constexpr float half_len = 0.5f;
    constexpr float pole_mass_len = mass_pole * half_len;
    constexpr float force_mag = 10.f;
    constexpr float tau = 0.02f;
    constexpr float theta_limit = 0.20943951f;   // 12 degrees
    constexpr float x_limit = 2.4f;

    float force = action.push == 1 ? force_mag : -force_mag;
    float cos_t = polyCos(s.theta);
    float sin_t = polySin(s.theta);

    float temp = (force + pole_mass_len * s.thetaDot * s.thetaDot * sin_t) /
        total_mass;
    float theta_acc = (gravity * sin_t - cos_t * temp) /
        (half_len * (4.f / 3.f - mass_pole * cos_t * cos_t / total_mass));
    float x_acc = temp - pole_mass_len * theta_acc * cos_t / total_mass;

    s.x = s.x + tau * s.xDot;
    s.xDot = s.xDot + tau * x_acc;
    s.theta = s.theta + tau * s.thetaDot;
    s.thetaDot = s.thetaDot + tau * theta_acc;
    count.t += 1;

    bool failed = s.x < -x_limit || s.x > x_limit ||
        s.theta < -theta_limit || s.theta > theta_limit;
    bool timeout = count.t >= ctx.data().maxSteps;

    reward.v = failed ? 0.f : 1.f;
    done.v = (failed || timeout) ? 1 : 0;
}

void Sim::setupTasks(TaskGraphManager &mgr, const Config &)
{
    TaskGraphBuilder &builder = mgr.init(TaskGraphID::Step);
    builder.addToGraph<ParallelForNode<Engine, stepSystem,
        Action, CartState, Reward, Done, StepCount>>({});
}

Sim::Sim(Engine &ctx, const Config &cfg, const WorldInit &init)
    : WorldBase(ctx),
      rng(init.seed),
      maxSteps(cfg.maxSteps),
      episode(0)
{
    cart = ctx.makeEntity<Cart>();
    ctx.get<Action>(cart).push = 0;
    ctx.get<Reward>(cart).v = 0.f;
    ctx.get<Done>(cart).v = 0;
    resetCart(ctx, ctx.get<CartState>(cart), ctx.get<StepCount>(cart));
    ctx.singleton<WorldReset>().reset = 0;
}

}

#ifdef MADRONA_GPU_MODE
MADRONA_BUILD_MWGPU_ENTRY(cartpole::Engine, cartpole::Sim,
                          cartpole::Config, cartpole::WorldInit);
#endif
