"""AlonaBase — the parameter surface of the drop-in (mirror of alona/alonabase.py:29-85 and of the CLI defaults of
alona/alona.py:24-120 for the flags the accelerated path reads).

`set_params` fills the reference's defaults, applies the reference's validations (each failing one goes through
log_error -> sys.exit(1) like alonabase.py:67-81) and rejects the options this package does not implement.
"""
import random
import string

from .log import log_debug, log_error, log_warning

# flag -> default (alona/alona.py: click options)
DEFAULTS = {
    'output_directory': None, 'dataformat': 'raw', 'minreads': 1000, 'minexpgenes': 0.001, 'qc_auto': True,
    'mrnafull': False, 'exclude_gene': None, 'remove_mito': 'no', 'hvg_method': 'seurat', 'hvg_n': 1000,
    'pca': 'irlb', 'pca_n': 75, 'nn_k': 10, 'prune_snn': 0.067, 'leiden_partition': 'RBERVertexPartition',
    'leiden_res': 0.8, 'ignore_small_clusters': 10, 'annotations': None, 'custom_clustering': None,
    'embedding': 'tSNE', 'perplexity': 30, 'species': 'mouse', 'loglevel': 'regular', 'seed': None,
}
HVG_METHODS = ('seurat', 'Brennecke2013', 'scran', 'Chen2016', 'M3Drop_smartseq2', 'M3Drop_UMI')


def random_str(length=8):
    return ''.join(random.choice(string.ascii_lowercase) for _ in range(length))


class AlonaBase:
    """AlonaBase class"""

    def __init__(self):
        self.params = getattr(self, 'params', None)

    def set_params(self, params):
        """ Set initial parameters (alona/alonabase.py:37-81). """
        if params is None:
            params = {}
        unknown = [k for k in params if k not in DEFAULTS]
        self.params = dict(DEFAULTS)
        self.params.update(params)
        self.anno = None
        if self.get_wd() is None:
            self.params['output_directory'] = 'alona_out_%s' % random_str()
        if self.params['loglevel'] == 'debug':
            log_debug('*** parameters *********************')
            for par in self.params:
                log_debug('%s : %s' % (par, self.params[par]))
            log_debug('************************************')
        for k in unknown:
            log_debug('parameter "%s" is not read by the accelerated path' % k)
        # the reference's validations (alonabase.py:67-81)
        if self.params['minreads'] < 0:
            log_error('--minreads must be a positive integer.')
        if self.params['minexpgenes'] < 0:
            log_error('--minexpgenes must be a positive value.')
        if self.params['nn_k'] < 0:
            log_error('--nn_k cannot be negative.')
        if self.params['prune_snn'] < 0 or self.params['prune_snn'] > 1:
            log_error('--prune_snn must have a value within [0,1]')
        if self.params['perplexity'] < 0:
            log_error('Perplexity cannot be less than 0.')
        if self.params['perplexity'] > 100:
            log_warning('Recommended values of --perplexity is 5-50.')
        if self.params['hvg_n'] <= 0:
            log_error('--hvg must be a positive value')
        # click.Choice checks of the CLI (alona/alona.py:53-61)
        if self.params['hvg_method'] not in HVG_METHODS:
            log_error('Unknown hvg method specified.')
        if self.params['pca'] not in ('irlb', 'regular'):
            log_error('--pca must be irlb or regular.')
        if self.params['pca_n'] <= 0:
            log_error('--pca_n must be a positive value')

    def get_wd(self):
        """ Retrieves the name of the output directory. """
        return self.params['output_directory']
