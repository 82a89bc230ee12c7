"""Parameter sets of the variogram goldens (reference test/test_variogram.py:22-110 plus one general-nu Matern);
shared by make_golden.py and the tests."""

VARIOGRAM_CASES = [
    ("GaussianVariogram", {"psill": 1.2, "nugget": 0, "range": 1200}),
    ("GaussianVariogram", {"psill": 1.2, "nugget": 0.101, "range": 1200}),
    ("ExponentialVariogram", {"psill": 1.2, "nugget": 0, "range": 1200}),
    ("SphericalVariogram", {"psill": 1.2, "nugget": 0, "range": 1200}),
    ("MaternVariogram", {"psill": 1.2, "nugget": 0, "range": 1200, "nu": 1.5, "method": "sklearn"}),
    ("MaternVariogram", {"psill": 1.2, "nugget": 0, "range": 1200, "nu": 0.5, "method": "sklearn"}),
    ("MaternVariogram", {"psill": 1.2, "nugget": 0, "range": 1200, "nu": 3.0, "method": "sklearn"}),
    ("MaternVariogram", {"psill": 1.2, "nugget": 0, "range": 1200, "nu": 1.5, "method": "gstat"}),
    ("MaternVariogram", {"psill": 1.2, "nugget": 0, "range": 1200, "nu": 1.5, "method": "karspeck"}),
    ("MaternVariogram", {"psill": 1.2, "nugget": 0.105, "range": 1200, "nu": 1.5, "method": "karspeck"}),
    ("MaternVariogram", {"psill": 0.9, "nugget": 0.02, "effective_range": 3000, "nu": 2.2, "method": "sklearn"}),
]
