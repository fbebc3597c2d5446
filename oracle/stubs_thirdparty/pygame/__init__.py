"""Import stub for the reference joystick script. Oracle tooling only."""
