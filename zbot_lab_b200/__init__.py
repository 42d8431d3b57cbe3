"""zbot_lab_b200 -- B200-native batched environment step for the ZBOT ``zbot-6b-walking-v2`` task.

The compute path is ``zbot_lab_b200/csrc`` (hand-written CUDA for sm_100a behind the C ABI in
``include/zbot_b200.h``); everything else here is the host-side mirror of the reference's task /
env / wrapper interface.  Importing ``zbot_lab_b200.tasks`` registers the gym ids.
"""
__version__ = "0.1.0"
