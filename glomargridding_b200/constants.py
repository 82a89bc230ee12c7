"""Constants shared with the reference (`glomar_gridding/constants.py:18,26-27`)."""

RADIUS_OF_EARTH_KM: float = 6371.0
KM_TO_NM: float = 1.852
NM_PER_LAT: float = 60.0
DEFAULT_N_JOBS: int = 4
DEFAULT_BACKEND: str = "loky"
