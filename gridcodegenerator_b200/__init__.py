"""B200-native rigid-body dynamics behind the GRiDCodeGenerator API."""
from .robot import Robot
from .facade import GRiDCodeGenerator
from .urdf import load_urdf, load_named_robot, parse_urdf_string, NAMED_ROBOTS

__all__ = ["GRiDCodeGenerator", "Robot", "load_urdf", "load_named_robot", "parse_urdf_string", "NAMED_ROBOTS"]
