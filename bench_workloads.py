"""Physics workloads for bench.py (kept separate so bench.py imports even when a
checkout lacks the physics fixture)."""

EXTRA_WORKLOADS = {
    # BASELINE.json configs[1]: "Madrona Escape Room 8192 worlds/GPU, physics+no-render, 1xB200"
    "room": dict(sim="room", worlds=8192, cfg={"episode_len": 200, "seed": 0},
                 ref_worlds=512, ref_steps=3000, taskgraphs=[0],
                 desc="rigid-body room (Escape-Room-class fixture, BASELINE configs[1]): 33 bodies/world "
                      "(2 agents, 15 cubes, 15 static hulls, plane), BVH broadphase + SAT narrowphase + "
                      "XPBD 4 substeps dt=0.04, 16-ray lidar, episode reset every 200 steps "
                      "(destroy+recreate cubes), no render"),
}
EXTRA_WORKLOADS["sortcheck"] = dict(
    sim="sortcheck", worlds=65536, cfg={"items_per_world": 48, "key_mask": 0xFFFFFFFF, "seed": 0},
    ref_worlds=0, taskgraphs=[0],
    desc="ECS-sort microbenchmark: 65536 worlds x 48 rows (3.1M rows, 38 B/row), every step re-keys all rows "
         "and runs SortArchetypeNode<Item, SortKey> (4 radix passes + fused 5-column permutation)")
# BASELINE.json configs[3]: "Escape Room 16384 worlds/GPU with 64x64 batch ray-traced observations"
EXTRA_WORKLOADS["room_render"] = dict(
    sim="room_render", worlds=16384,
    cfg={"episode_len": 200, "seed": 0, "resolution": 64, "rgbd": True},
    ref_worlds=0, taskgraphs=[0], render=True,
    desc="rigid-body room + 2 cameras/world, 64x64 RGBA8 + f32 depth by the batch ray caster "
         "(BASELINE configs[3] class; GPU only: the reference CPU backend cannot ray cast)")
# BASELINE.json configs[2]: "Hide&Seek 4096 worlds/GPU, full XPBD rigid-body + BVH, 1->8 B200"
EXTRA_WORKLOADS["arena"] = dict(
    sim="arena", worlds=4096, cfg={"episode_len": 200, "seed": 0},
    ref_worlds=512, ref_steps=1500, taskgraphs=[0],
    desc="rigid-body arena (Hide&Seek-class fixture, BASELINE configs[2]): 49 bodies/world (6 hexagonal-prism "
         "agents, 12 cubes, 8 long boxes, 3 wedge ramps, 2 barrels, 2 latched doors, 15 static hulls, plane), "
         "fixed joints (latches, grab) + one-step hinge joints (shove), BVH broadphase + SAT narrowphase + XPBD "
         "4 substeps dt=0.04, 16-ray lidar + 5 line-of-sight rays per agent, episode reset every 200 steps "
         "(27 bodies + joints destroyed and recreated), no render")
EXTRA_DEFAULT = "room"
