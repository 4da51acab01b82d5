"""Output artefact names of the hot path — same relative paths as the reference writes
(alona/constants.py:17-39), because the unchanged downstream steps and the cache-skip rules
(alona/cell.py:425-430, alona/clustering.py:197-201) key on them."""

OUTPUT = {
    'FILENAME_PCA': '/csvs/pca.csv',
    'FILENAME_EMBEDDING_PREFIX': '/csvs/embeddings_',
    'FILENAME_HVG': '/csvs/highly_variable_genes.tsv',
    'FILENAME_SNN_GRAPH': '/csvs/snn_graph.csv',
    'FILENAME_CLUSTERS_LEIDEN': '/csvs/clusters_leiden.csv',
    'FILENAME_MEDIAN_EXP': '/csvs/median_exp.tsv',
    'FILENAME_MEAN_EXP': '/csvs/mean_exp.tsv',
    'FILENAME_SETTINGS': '/settings.txt',
    'FILENAME_KNN_map': '/KNN.joblib',
}
