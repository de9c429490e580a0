"""Drop-in boundary, end to end on the HOST side (north_star "interfaces/super_resolution.py and main.py run unchanged"):
with ``tpgsr_b200/shadow`` in front of an UNMODIFIED reference checkout on ``sys.path``, ``import interfaces.super_resolution`` succeeds,
the hot-path names bind to the tpgsr_b200 drop-ins, every other submodule still loads from the reference's own files, and the reference's
own ``TextBase.generator_init`` / ``CRNN_init`` construct the drop-ins with the reference's kwargs and print the complexity through the
shadowed ``ptflops``.  Runs only where a reference checkout exists (this container: /root/reference; never on the GPU box), in a child
process so that the shadowed top-level names (``model``, ``loss``, ``utils``) do not leak into the test session.

Third-party packages the reference imports but this image lacks are stubbed (they are outside the hot path): IPython, easydict, thop,
tensorboardX, lmdb, pyfasttext, editdistance."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('TPGSR_REFERENCE_ROOT', '/root/reference')

CHILD = textwrap.dedent(r'''
    import contextlib, os, sys, types
    ROOT, REF = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ROOT)
    import tpgsr_b200.shadow as shadow
    sys.path[:0] = [shadow.path(), REF]

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = EasyDict(v) if isinstance(v, dict) else v
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    # third-party packages the reference imports but this image lacks (all outside the hot path): stubbed on demand, submodules included
    STUBS = {'IPython': dict(embed=lambda *a, **k: None), 'easydict': dict(EasyDict=EasyDict), 'thop': dict(profile=lambda *a, **k: (0, 0)),
             'tensorboardX': dict(SummaryWriter=object), 'lmdb': {}, 'pyfasttext': dict(FastText=object), 'editdistance': dict(eval=lambda a, b: 0),
             'matplotlib': {}, 'mpl_toolkits': {}, 'imgaug': {}, 'hdf5storage': {}, 'skimage': {}}
    import importlib.abc, importlib.machinery

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return type(name, (), {})

    class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path=None, target=None):
            top = fullname.split('.')[0]
            if top in STUBS:
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None
        def create_module(self, spec):
            m = _Anything(spec.name)
            m.__path__ = []
            if spec.name in STUBS:
                m.__dict__.update(STUBS[spec.name])
            return m
        def exec_module(self, module):
            pass
    for name in list(STUBS):
        try:
            __import__(name)
            del STUBS[name]                       # really installed: use it
        except ImportError:
            pass
    sys.meta_path.append(StubFinder())
    # the reference's dataset code calls scipy aliases that newer scipy releases dropped (utils/utils_deblur.py:494 runs at import time)
    import numpy, scipy
    for alias in ('finfo', 'pi', 'exp', 'sqrt'):
        if not hasattr(scipy, alias):
            setattr(scipy, alias, getattr(numpy, alias))

    import torch
    import interfaces.super_resolution as sr            # the reference's file, unmodified
    from interfaces import base
    import tpgsr_b200.model.tsrn, tpgsr_b200.model.crnn, tpgsr_b200.model.srcnn, tpgsr_b200.model.srresnet
    import tpgsr_b200.loss.image_loss, tpgsr_b200.loss.semantic_loss, tpgsr_b200.utils.ssim_psnr, tpgsr_b200.utils.metrics

    assert os.path.samefile(os.path.dirname(sr.__file__), os.path.join(REF, 'interfaces'))
    # hot-path names -> the drop-ins
    assert base.tsrn.TSRN is tpgsr_b200.model.tsrn.TSRN and base.tsrn.TSRN_TL is tpgsr_b200.model.tsrn.TSRN_TL
    assert base.crnn.CRNN is tpgsr_b200.model.crnn.CRNN
    assert base.srcnn.SRCNN is tpgsr_b200.model.srcnn.SRCNN and base.srresnet.SRResNet_TL is tpgsr_b200.model.srresnet.SRResNet_TL
    assert base.image_loss.ImageLoss is tpgsr_b200.loss.image_loss.ImageLoss
    assert type(sr.sem_loss) is tpgsr_b200.loss.semantic_loss.SemanticLoss
    assert base.ssim_psnr.calculate_psnr is tpgsr_b200.utils.ssim_psnr.calculate_psnr and type(sr.ssim) is tpgsr_b200.utils.ssim_psnr.SSIM
    assert sr.get_string_crnn is tpgsr_b200.utils.metrics.get_string_crnn
    assert sr.str_filt('Ab-3!', 'lower') == 'ab3'
    # everything else -> the reference's own files
    for mod in (base.bicubic, base.vdsr, base.edsr, base.rdn, base.lapsrn, base.esrgan, base.recognizer, base.moran, base.gradient_loss,
                base.percptual_loss, base.utils_moran, base.utils_crnn, sr.utils_moran, sr.gumbel_softmax):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REF) + os.sep), mod.__file__
    assert sr.get_string_aster.__module__ == 'utils._reference_metrics' and callable(sr.Accuracy)
    assert callable(base.get_vocabulary)
    try:
        base.tsrn.TSRN_C2F
    except AttributeError as e:
        assert 'out of scope' in str(e)
    else:
        raise AssertionError('out-of-scope names must raise')

    # the reference's own constructors over the drop-ins (CPU: construction + complexity printout only — compute needs the GPU)
    torch.cuda.device = lambda idx: contextlib.nullcontext()
    cfg = EasyDict(TRAIN=dict(width=128, height=32, ngpu=1, down_sample_scale=2, voc_type='all', VAL=dict(crnn_pretrained='/nonexistent')))
    for arch, cls, crit in (('tsrn', base.tsrn.TSRN, base.image_loss.ImageLoss), ('tsrn_tl_cascade', base.tsrn.TSRN_TL, base.image_loss.ImageLoss),
                            ('srcnn', base.srcnn.SRCNN, torch.nn.MSELoss), ('srresnet_tl', base.srresnet.SRResNet_TL, torch.nn.MSELoss)):
        obj = object.__new__(sr.TextSR)
        obj.config, obj.scale_factor, obj.mask, obj.device, obj.resume = cfg, 2, True, torch.device('cpu'), ''
        obj.args = EasyDict(arch=arch, STN=arch.startswith('tsrn'), srb=5, hd_u=32, gradient=True, test=False)
        out = obj.generator_init()
        assert type(out['model']) is cls and type(out['crit']) is crit, (arch, type(out['model']))
        assert out['model'].module is out['model']
    obj = object.__new__(sr.TextSR)
    obj.config, obj.device = cfg, torch.device('cpu')
    try:
        obj.CRNN_init()
    except FileNotFoundError:
        pass                                     # constructed the drop-in CRNN, printed its complexity, then looked for crnn.pth
    else:
        raise AssertionError('expected the missing checkpoint to be reported')
    print('SHADOW-OK')
''')


@pytest.mark.timeout(600)
def test_reference_interfaces_import_and_construct_over_shadow_packages(tmp_path):
    if not os.path.isfile(os.path.join(REF, 'interfaces', 'super_resolution.py')):
        pytest.skip('no reference checkout here (the GPU box has none by design)')
    script = tmp_path / 'child.py'
    script.write_text(CHILD)
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    r = subprocess.run([sys.executable, str(script), ROOT, REF], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=550)
    assert r.returncode == 0 and 'SHADOW-OK' in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
    assert 'Computational complexity' in r.stdout and 'MMac' in r.stdout and 'TSRN_TL: 959.24 MMac' in r.stdout


def test_shadow_ptflops_counts_match_survey():
    """The analytic complexity of the drop-ins equals the hooked forward FLOPs SURVEY §8d lists for the reference modules."""
    sys.path.insert(0, os.path.join(ROOT, 'tpgsr_b200', 'shadow'))
    try:
        import importlib
        ptf = importlib.import_module('ptflops')
        assert ptf.__file__.startswith(os.path.join(ROOT, 'tpgsr_b200', 'shadow'))
        from tpgsr_b200.model.crnn import CRNN
        from tpgsr_b200.model.srcnn import SRCNN
        from tpgsr_b200.model.srresnet import SRResNet_TL
        from tpgsr_b200.model.tsrn import TSRN, TSRN_TL
        kw = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
        for m, res, gflop in ((TSRN(**kw), (4, 16, 64), 1.8062), (TSRN_TL(**kw), (4, 16, 64), 1.9185), (CRNN(32, 1, 37, 256), (1, 32, 100), 1.4103),
                              (SRCNN(scale_factor=2, width=128, height=32, STN=False), (3, 16, 64), 0.1638),
                              (SRResNet_TL(scale_factor=2, width=128, height=32, STN=False, mask=True), (4, 16, 64), 1.9425)):
            macs, params = ptf.get_model_complexity_info(m, res, as_strings=False, print_per_layer_stat=False)
            assert abs(2 * macs / 1e9 - gflop) < 1e-4, (type(m).__name__, 2 * macs / 1e9)
            assert params == sum(p.numel() for p in m.parameters())
            s_macs, s_params = ptf.get_model_complexity_info(m, res, as_strings=True, print_per_layer_stat=False)
            assert s_macs.endswith('Mac') and s_params.endswith(('M', 'k'))
        # a module without complexity_info is counted the ptflops way (hooks on a CPU forward)
        import torch
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1))
        macs, _ = ptf.get_model_complexity_info(net, (3, 5, 7), as_strings=False, print_per_layer_stat=False)
        assert macs == 3 * 8 * 9 * 35 + 8 * 4 * 35
    finally:
        sys.path.remove(os.path.join(ROOT, 'tpgsr_b200', 'shadow'))
        for k in [k for k in sys.modules if k == 'ptflops' or k.startswith('ptflops.')]:
            del sys.modules[k]
