"""The config/path glue of the reference's utils/__init__.py that the hot path needs
(get_cachedir :28-31, get_downsampling :47-49, calc_cell_width_height :52-56, load_config :69-73)."""
import importlib
import os


def get_cachedir(config):
    basedir = os.path.expanduser(os.path.expandvars(config.get('config', 'basedir')))
    name = os.path.basename(config.get('cache', 'names'))
    return os.path.join(basedir, 'cache', name)


def get_downsampling(config):
    model = config.get('config', 'model')
    mod = importlib.import_module('.'.join(['yolo_tf_b200', 'model', model, 'inference']))
    return getattr(mod, config.get(model, 'inference').upper() + '_DOWNSAMPLING')


def calc_cell_width_height(config, width, height):
    downsampling_width, downsampling_height = get_downsampling(config)
    assert width % downsampling_width == 0
    assert height % downsampling_height == 0
    return width // downsampling_width, height // downsampling_height


def load_config(config, paths):
    for path in paths:
        path = os.path.expanduser(os.path.expandvars(path))
        assert os.path.exists(path)
        config.read(path)
