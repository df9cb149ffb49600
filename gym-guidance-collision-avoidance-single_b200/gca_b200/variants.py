"""Variant table: one row per environment class of the reference -> a gca_config.

Rows restate the reward / termination table of SURVEY.md 8(a); every constant cites the
reference line it comes from (PKG = gym_guidance_collision_avoidance_single/envs).
"""
from . import abi

# name: (action_kind, obs_kind, wall_kind, shaped_default, (r_nmac, r_conflict, r_wall, r_goal, r_default), uses max_steps)
VARIANTS = {
    # PKG/SingleAircraftEnv.py:130-133 (discrete 9), :170 NMAC -20, :174 conflict -5, :182 goal +10, :183 -d/1200
    "SingleAircraftEnv": (abi.ACT_DISCRETE9, abi.OBS_VECTOR, abi.WALL_NONE, 1, (-20, -5, 0, 10, 0), False),
    # PKG/SingleAircraft2Env.py:163 NMAC -5, :167 conflict -1, :169-170 wall -100 terminal, :174 goal +1
    "SingleAircraft2Env": (abi.ACT_CONTINUOUS2, abi.OBS_VECTOR, abi.WALL_TERMINAL, 1, (-5, -1, -100, 1, 0), False),
    # PKG/SingleAircraftHEREnv.py:179 NMAC -5, :183 conflict -1, :191 goal 0, :192 default -1
    "SingleAircraftHEREnv": (abi.ACT_CONTINUOUS2, abi.OBS_HER, abi.WALL_NONE, 0, (-5, -1, 0, 0, -1), False),
    # PKG/SingleAircraftDiscreteHEREnv.py:170 NMAC -5, :174 conflict -1, :176-177 wall -5 terminal, :181 goal +1, :182 default 0
    "SingleAircraftDiscreteHEREnv": (abi.ACT_DISCRETE3, abi.OBS_DHER, abi.WALL_TERMINAL, 0, (-5, -1, -5, 1, 0), False),
    # PKG/SingleAircraftStackEnv.py:134-136 max steps, :163 NMAC -20, :167 conflict -5, :170-171 wall -10 non-terminal, :175 goal +10000
    "SingleAircraftStackEnv": (abi.ACT_DISCRETE9, abi.OBS_NONE, abi.WALL_PENALTY, 1, (-20, -5, -10, 10000, 0), True),
    # Simulators/SingleAircraftMCTSEnv.py:163-180: rewards come from Simulators/config.py:37-43 (filled by make_config)
    "SingleAircraftMCTSEnv": (abi.ACT_DISCRETE9, abi.OBS_RAW, abi.WALL_NONE, 1, None, False),
    # Simulators/SingleAircraftMCTSRandIntruderEnv.py: the MCTS env whose intruders turn at random after every step
    # (_update_headings :166-174), drift by Config.position_sigma per step (:183) and show (speed, heading) in the raw
    # observation (:133-140); reward row from Simulators/config.py:37-43 (:207-227); info is the bare string (:164)
    "SingleAircraftMCTSRandIntruderEnv": (abi.ACT_DISCRETE9, abi.OBS_RAW6, abi.WALL_NONE, 1, None, False),
    # Simulators/SingleAircraftDiscrete9HEREnv.py: random ownship start :78-82, n nearest intruders :127-144, rewards
    # from Simulators/config.py:37-43 (:207-227), out-of-map rule only under sparse_reward (:217-219)
    "SingleAircraftDiscrete9HEREnv": (abi.ACT_DISCRETE9, abi.OBS_NEAREST, None, 1, None, False),
    # Simulators/SingleAircraftDiscrete3HEREnv.py: as 9HER with Discrete(3) heading-only actions (:407-411), the goal drawn
    # 100 px inside the map (:349-353) and the nearest-intruder term added to the default reward (:225-232)
    "SingleAircraftDiscrete3HEREnv": (abi.ACT_DISCRETE3_HEADING, abi.OBS_NEAREST, None, 1, None, False),
    # Simulators/SingleAircraftEnv.py ("deprecated" copy of the registered env): Config-driven reward row (:168-185), no
    # out-of-map rule (:175-176 commented), info dict (:139); SingleAircraftRandomEnv.py: the same with a random start (:71-73)
    "SimSingleAircraftEnv": (abi.ACT_DISCRETE9, abi.OBS_VECTOR, abi.WALL_NONE, 1, None, False),
    "SingleAircraftRandomEnv": (abi.ACT_DISCRETE9, abi.OBS_VECTOR, abi.WALL_NONE, 1, None, False),
}


def default_config_class(variant):
    """The Config class the reference's class of this name reads when none is given: the Simulators/ copies (Config-
    driven reward row `rewards=None`, or the nearest-n observation) `import config` = Simulators/config.py (it has
    NMAC_penalty / sparse_reward / n / diagonal ...), the registered classes the package's envs/config.py."""
    row = VARIANTS[variant]
    if row[4] is None or row[1] == abi.OBS_NEAREST:
        from Simulators.config import Config
    else:
        from gym_guidance_collision_avoidance_single.envs.config import Config
    return Config


def make_config(variant, cfg_cls, time_limit=0):
    """Snapshot the class attributes of `cfg_cls` (the reference's Config idiom,
    PKG/SingleAircraftEnv.py:49-64, :286-297) into a gca_config for `variant`."""
    act, obs, wall, shaped, rewards, use_max = VARIANTS[variant]
    c = abi.GcaConfig()
    c.window_width = cfg_cls.window_width
    c.window_height = cfg_cls.window_height
    c.minimum_separation = cfg_cls.minimum_separation
    c.nmac_dist = cfg_cls.NMAC_dist
    c.initial_min_dist = cfg_cls.initial_min_dist
    c.goal_radius = cfg_cls.goal_radius
    c.min_speed = cfg_cls.min_speed
    c.max_speed = cfg_cls.max_speed
    c.d_speed = cfg_cls.d_speed
    c.speed_sigma = cfg_cls.speed_sigma
    c.d_heading = cfg_cls.d_heading
    c.heading_sigma = cfg_cls.heading_sigma
    refresh_observation_params(c, cfg_cls)
    if rewards is None:   # Config-driven reward row of the MCTS env
        rewards = (cfg_cls.NMAC_penalty, cfg_cls.conflict_penalty, cfg_cls.wall_penalty, cfg_cls.goal_reward,
                   cfg_cls.step_penalty)
        shaped = 0 if cfg_cls.sparse_reward else 1
    if wall is None:      # `if Config.sparse_reward: if not position_range.contains(drone.position): ... 'w'`
        wall = abi.WALL_TERMINAL if cfg_cls.sparse_reward else abi.WALL_NONE
    c.r_nmac, c.r_conflict, c.r_wall, c.r_goal, c.r_default = [float(r) for r in rewards]
    c.shaped_default = shaped
    c.action_kind, c.obs_kind, c.wall_kind = act, obs, wall
    c.max_steps = int(cfg_cls.max_steps) if use_max else 0
    c.time_limit = int(time_limit)
    c.random_start = 0
    c.nearest_n = 0
    c.ob_diagonal = 1.0
    c.conflict_coeff = 0.0
    c.goal_margin = 0.0
    c.shaped_nearest = 0
    c.intruder_turns = 0
    c.position_drift = 0.0
    c.turn_prob = 0.0
    c.turn_max_deg = 0.0
    if variant == "SingleAircraftMCTSRandIntruderEnv":
        c.intruder_turns = 1
        c.position_drift = float(cfg_cls.position_sigma)     # :82, :183
        c.turn_prob = 0.1                                    # :170
        c.turn_max_deg = 10.0                                # :173
    if obs == abi.OBS_NEAREST:
        c.random_start = 1
        c.nearest_n = int(cfg_cls.n)
        c.ob_diagonal = cfg_cls.diagonal
    if variant == "SingleAircraftRandomEnv":
        c.random_start = 1
    if variant == "SingleAircraftDiscrete3HEREnv":
        c.conflict_coeff = cfg_cls.conflict_coeff
        c.goal_margin = 100.0
        c.shaped_nearest = 1
    return c


def refresh_observation_params(c, cfg_cls):
    """_get_ob reads these from the Config class on every call, not from the instance (Q12)."""
    c.ob_window_width = cfg_cls.window_width
    c.ob_window_height = cfg_cls.window_height
    c.ob_min_speed = cfg_cls.min_speed
    c.ob_max_speed = cfg_cls.max_speed
    if getattr(c, "obs_kind", None) == abi.OBS_NEAREST:
        c.ob_diagonal = cfg_cls.diagonal      # Config.diagonal is read inside _get_ob as well


def obs_dim(c, n):
    if c.obs_kind in (abi.OBS_VECTOR, abi.OBS_RAW):
        return 4 * n + 8
    if c.obs_kind in (abi.OBS_HER, abi.OBS_DHER):
        return 4 * n + 6
    if c.obs_kind == abi.OBS_NEAREST:
        return 4 + 5 * c.nearest_n
    if c.obs_kind == abi.OBS_RAW6:
        return 6 * n + 8
    return 0


# reward values the reference returns as Python ints (Q11): the single-env facade restores the type
def reward_is_int(variant, info_code):
    row = VARIANTS[variant][4]
    if row is None or info_code == abi.INFO_NONE and VARIANTS[variant][3]:
        return False
    return True
