"""CPU oracle for the 2048-PPO hot path -- TEST INFRASTRUCTURE ONLY (see oracle2048.c)."""
