"""CPU tier: the parameter surface of the drop-in (alona/alonabase.py:37-81, CLI defaults of alona/alona.py)."""
import pytest

from alona_b200.alonabase import AlonaBase, DEFAULTS


def test_defaults_are_the_cli_defaults():
    o = AlonaBase()
    o.set_params({'output_directory': '/tmp/x'})
    # alona/alona.py: --hvg seurat, --hvg_n 1000, --pca irlb, --pca_n 75, --nn_k 10, --prune_snn 0.067,
    # --leiden_res 0.8, --ignore_small_clusters 10, --minreads 1000, --minexpgenes 0.001
    for key, val in (('hvg_method', 'seurat'), ('hvg_n', 1000), ('pca', 'irlb'), ('pca_n', 75), ('nn_k', 10),
                     ('prune_snn', 0.067), ('leiden_res', 0.8), ('ignore_small_clusters', 10), ('minreads', 1000),
                     ('minexpgenes', 0.001), ('remove_mito', 'no'), ('embedding', 'tSNE')):
        assert o.params[key] == val, key
    assert o.get_wd() == '/tmp/x'
    assert set(DEFAULTS) <= set(o.params)


def test_missing_output_directory_gets_a_random_name():
    o = AlonaBase()
    o.set_params(None)
    assert o.get_wd().startswith('alona_out_') and len(o.get_wd()) == len('alona_out_') + 8


@pytest.mark.parametrize('bad', [{'minreads': -1}, {'minexpgenes': -0.5}, {'nn_k': -3}, {'prune_snn': 1.5},
                                 {'prune_snn': -0.1}, {'perplexity': -1}, {'hvg_n': 0}, {'hvg_method': 'nonsense'},
                                 {'pca': 'sparse'}, {'pca_n': 0}])
def test_invalid_values_exit_like_log_error(bad):
    o = AlonaBase()
    with pytest.raises(SystemExit) as e:
        o.set_params(dict(bad, output_directory='/tmp/x'))
    assert e.value.code == 1                              # alona/log.py:52-54
