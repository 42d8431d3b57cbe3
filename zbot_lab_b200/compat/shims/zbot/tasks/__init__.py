import zbot_lab_b200.tasks  # noqa: F401  (gym.register("zbot-6b-walking-v2", ...))
