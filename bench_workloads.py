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
EXTRA_DEFAULT = "room"
