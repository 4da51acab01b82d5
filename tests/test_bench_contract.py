"""CPU tier: the arithmetic behind bench.py's `roofline` / `roofline_stages` objects (no GPU, no timing)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location('bench_module', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


WL = {'cells': 1306127, 'genes': 27998, 'hvg_n': 2000, 'pca_n': 50, 'k': 20}


def test_roofline_object_of_the_dominant_kernel():
    b = _bench()
    r = b.roofline(WL, 1, {'knn_main': 500.0}, {'bf16_tflops_sustained': 1395.5, 'hbm_gbs': 6538.9})
    assert set(['bound', 'achieved', 'peak', 'unit', 'frac', 'traffic']).issubset(r)
    assert r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s' and r['peak'] == 1395.5
    flops = 2.0 * WL['cells'] * WL['cells'] * WL['pca_n']                 # 2.N^2.d, true d
    assert abs(r['achieved'] - flops / 0.5 / 1e12) < 1e-6 and abs(r['frac'] - r['achieved'] / 1395.5) < 1e-12
    assert r['frac'] < 50 / 112                                           # two FP16 products + K padding cap it
    # sharded: each rank answers N/world queries against all N points
    r8 = b.roofline(WL, 8, {'knn_main': 500.0 / 8}, {'bf16_tflops_sustained': 1395.5})
    assert abs(r8['achieved'] - r['achieved']) < 1e-6
    assert b.roofline(WL, 1, {}, {}) is None                              # no event pair -> no claim


def test_stage_rooflines_use_the_survey_byte_counts():
    b = _bench()
    info = {'nnz': 2408546910, 'nnz_hvg': 220675500, 'nmult': 518, 'sum_panel': 15526}
    out = {'sum_walk': 2487424504, 'edges': 27191773}
    fine = {'norm_moments': (29.0, 1), 'hvg_extract': (8.0, 2), 'irlb_atw': (105.0, 260), 'irlb_av': (179.0, 259),
            'irlb_panel': (75.0, 259), 'snn_count': (69.0, 1), 'snn_fill': (0.3, 1)}
    rows = {r['kernel']: r for r in b.stage_rooflines(WL, 1, fine, info, out, {'hbm_gbs': 6538.9})}
    k1 = rows['norm_moments_kernel']
    assert k1['bytes'] == info['nnz'] * 8 + (WL['cells'] + 1) * 8 + WL['cells'] * 8 + 2 * WL['genes'] * 8
    assert abs(k1['achieved'] - k1['bytes'] / 0.029 / 1e9) < 1e-6 and k1['peak'] == 6538.9
    av = rows['av_kernel (A v)']
    assert av['bytes'] == (518 / 2) * (info['nnz_hvg'] * 12 + WL['cells'] * 24)
    assert 'note' in rows['atw_compact_kernel (A^T w)']                   # algorithmic vs actually-read bytes is explained
    assert all(0 < r['frac'] for r in rows.values())
