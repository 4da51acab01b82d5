"""alona_b200 — B200-native clustering front end of alona (normalise -> HVG -> PCA -> kNN -> SNN).

Drop-in surface (same names as oscar-franzen/alona v0.3.5):
    alona_b200.cell.AlonaCell.normalize
    alona_b200.hvg.AlonaHighlyVariableGenes
    alona_b200.irlbpy.lanczos
    alona_b200.clustering.AlonaClustering  (find_variable_genes, PCA, knn, snn, cluster)
All arithmetic runs in libalona_b200.so (hand-written sm_100a CUDA, C-ABI in include/alona_b200.h).
"""
__version__ = '0.1.0'
