"""TEST INFRASTRUCTURE ONLY - CPU oracle for the node-feature assembly.

The notebook's own pandas code, statement for statement:
  * cell 21, /root/reference/hovernet_tile_inference.ipynb:2899-2909 (z-scores with mean / std(ddof=0));
  * cell 23, ipynb:2950-2957 (pd.get_dummies(final_df["type"], prefix="type"), onehot_cols + morph_z_cols).
pandas is the reference's dependency and is present here, so this IS the reference computation; what is
build-defined is only the final ``x = final_df[feat_cols]`` as float32 (the notebook never defines ``x``; its
printed shape [101, 15] is 5 one-hot + 10 z columns).  Pinned by tests/test_oracle.py on the notebook's stored
rows (Appendix D-2 fixtures) for the z-score rule.
"""
import numpy as np
import pandas as pd

CONT_COLS = ["area", "perimeter", "eccentricity", "solidity", "major_axis_length", "minor_axis_length",
             "perimeter_area", "compactness", "roundness", "elongation"]


def node_features(final_df: pd.DataFrame, cont_cols=CONT_COLS):
    final_df = final_df.copy()
    for col in cont_cols:                                   # cell 21
        if col in final_df.columns:
            mu = final_df[col].mean()
            sigma = final_df[col].std(ddof=0)
            if sigma == 0 or np.isnan(sigma):
                final_df[col + "_z"] = 0.0
            else:
                final_df[col + "_z"] = (final_df[col] - mu) / sigma
    type_onehot = pd.get_dummies(final_df["type"], prefix="type")   # cell 23
    onehot_cols = list(type_onehot.columns)
    final_df = pd.concat([final_df, type_onehot], axis=1)
    morph_z_cols = [c for c in final_df.columns if c.endswith("_z")]
    feat_cols = onehot_cols + morph_z_cols
    x = final_df[feat_cols].to_numpy(dtype=np.float64).astype(np.float32)
    return x, feat_cols
