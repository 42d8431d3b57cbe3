"""Static task parameters of ``zbot-6b-walking-v2`` (mirror of ``ZbotDirectEnvCfgV2``,
``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py:26-206``)."""

#: "train reward 2000 step4" (…env_v2.py:190-206); dict ORDER = evaluation order (SURVEY C-4)
REWARD_SCALES_V2 = {
    "base_vel_forward": 1.0,
    "feet_downward": -2.0,
    "feet_forward": -1.0,
    "base_heading_x": -1.0,
    "base_heading_x_sum": -5.0,
    "step_length": 5.0,
    "airtime_balance": -15.0,
    "action_rate": -0.1,
    "torques": -0.002,
    "feet_slide": -10.0,
    "base_pos_y_err": -2.0,
    "base_pos_y_err_sum": -2.0,
    "airtime_sum": 3.0,
}
