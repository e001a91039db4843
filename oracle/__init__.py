"""CPU oracle for the nuclei-table -> WSI space -> morphology + cell-graph hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``path_gene_multimodal_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may.  The product path has no CPU fallback.

What each module restates (reference file:line) and how it is pinned:

* ``tile_to_wsi``  - ``add_wsi_coords_to_nuclei`` (aggregated_hovernet_run.py:263-336).
  PINNED: the reference's own function is AST-extracted from /root/reference in the build
  container and run on seeded frames; its outputs are committed under ``tests/golden/``
  (generator: ``oracle/make_golden.py``) and the restatement is compared with them.
* ``graph.radius_graph`` - cells 23-26 of hovernet_tile_inference.ipynb (ipynb:2963-3042):
  ``scipy.spatial.cKDTree.query_ball_tree`` + the ``i < j`` filter + ``np.linalg.norm``.
  scipy is the reference's own (un-vendored, un-pinned) dependency; container has 1.18.1.
  PINNED by scipy itself + brute force; the notebook's stored count (1189 edges) has no
  recoverable input -> "parity unpinned by reference tests".
* ``graph.knn`` - cell 11 (ipynb:1815-1850): libpysal ``KNN.from_array`` == cKDTree.query(k+1)
  minus self (libpysal is absent here; restated), canonical ``(d^2, index)`` order.
  PINNED on the D-1 distances printed in the notebook output; neighbour sets unpinned.
* ``graph.undirected_union`` - cell 11 nx.Graph loop (ipynb:1879-1894), checked against networkx.
* ``morphology`` - shapely/GEOS ``area/length/centroid/bounds`` (polygon_morphology.py:240-248)
  and the cell-18 derived features (ipynb:2415-2456).  shapely / skimage are absent:
  float64 restatement.  Derived-feature formulas PINNED on the ten stored notebook rows
  (SURVEY Appendix D-2); polygon eccentricity vs the reference's raster eccentricity is
  "parity unpinned" by construction.
* ``graph.composition`` / ``degree_stats`` - not implemented in the reference (README.md:127,136);
  build-defined -> "parity unpinned", exact vs this oracle.
"""
