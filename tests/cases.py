"""Shared seeded cases for the parity tests (SURVEY.md section 8 config shapes,
scaled so the oracle finishes in seconds)."""
import numpy as np

from navsim import synthetic

# name -> dict(landscape kwargs, world kwargs, curve, start offset)
CASES = {
    # C1 family: sensor 40x2 @ 2x4 px, 10 headings over 180 deg, 5 levels
    "c1_small": dict(land=dict(seed=1001, side=600, sigma=6.0), curve=0.0, off=(0.1, 5.0),
                     world=dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=5.0,
                                n_test_angles=10, n_sensor_levels=5, max_distance_to_training_path=450)),
    # example.gif family: 8 levels, 20 headings over 60 deg
    "gif": dict(land=dict(seed=1002, side=500, sigma=3.0), curve=0.0, off=(0.0, 0.0),
                world=dict(sensor_dimensions=(40, 2), sensor_pixel_dimensions=(2, 4), step_size=2.0,
                           n_test_angles=20, n_sensor_levels=8, saccade_degrees=60.,
                           max_distance_to_training_path=450)),
    # chemistry on, centre mask, three hues
    "chem": dict(land=dict(seed=1003, side=600, sigma=6.0, n_chemicals=3), curve=0.5, off=(-0.3, -10.0),
                 world=dict(sensor_dimensions=(20, 4), sensor_pixel_dimensions=(2, 2), step_size=4.0,
                            n_test_angles=12, n_sensor_levels=8, mask_middle_n=2, chem_weight=0.3,
                            max_distance_to_training_path=450)),
    # square 1x1 sensor, full-circle sweep (first and last heading coincide), tight max distance
    "square": dict(land=dict(seed=1004, side=500, sigma=4.0), curve=0.5, off=(0.2, 10.0),
                   world=dict(sensor_dimensions=(16, 16), sensor_pixel_dimensions=(1, 1), step_size=6.0,
                              n_test_angles=20, n_sensor_levels=3, saccade_degrees=360.,
                              max_distance_to_training_path=60)),
    # tiny sensor: integer ties on every step
    "ties": dict(land=dict(seed=1005, side=500, sigma=6.0, n_chemicals=2), curve=0.0, off=(0.5, 20.0),
                 world=dict(sensor_dimensions=(8, 2), sensor_pixel_dimensions=(2, 2), step_size=2.0,
                            n_test_angles=10, n_sensor_levels=5, max_distance_to_training_path=450)),
    # pure chemistry metric
    "chem1": dict(land=dict(seed=1006, side=500, sigma=6.0, n_chemicals=2), curve=0.0, off=(0.2, -5.0),
                  world=dict(sensor_dimensions=(8, 2), sensor_pixel_dimensions=(2, 2), step_size=2.0,
                             n_test_angles=10, n_sensor_levels=5, chem_weight=1.0,
                             max_distance_to_training_path=450)),
}


def build_case(name):
    c = CASES[name]
    L = synthetic.make_landscape(kind="stitch", **c["land"])
    w = dict(c["world"])
    tpath = synthetic.training_path_for(L.shape, w["step_size"], w["n_test_angles"], c["curve"])
    spw = w["sensor_dimensions"][0] * w["sensor_pixel_dimensions"][0]
    pose = synthetic.start_pose(tpath, c["off"], spw)
    frames = synthetic.default_frames(tpath, w["step_size"])
    return L, w, tpath, pose, frames


def agent_grid(tpath, w, n_lat=3, n_deg=3):
    spw = w["sensor_dimensions"][0] * w["sensor_pixel_dimensions"][0]
    return synthetic.start_pose_grid(tpath, spw, n_lat=n_lat, n_deg=n_deg, lat=0.3, deg=10.0)
