"""A TensorFlow-1 API shim, just wide enough to EXECUTE the reference's own `lib/models.py` on torch-CPU.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in cape_b200/ imports it.

Why: the reference's arithmetic lives in tensorflow-gpu==1.13.2, which cannot be installed here, and the reference ships
no golden vectors -- so the hand-written restatement in oracle/cape_oracle.py had nothing of the reference's to be
checked against except the Laplacians.  But the reference's MODEL code is plain Python that only *calls* about seventy
TensorFlow symbols (`tf.reshape`, `tf.sparse_tensor_dense_matmul`, `tf.layers.dense`, `tf.variable_scope`,
`tf.train.MomentumOptimizer`, ...).  This module provides those symbols with their documented TF-1.13 semantics on
torch tensors, installs itself as `tensorflow` in `sys.modules`, and lets `lib.models.CAPE.build_graph()` run
UNMODIFIED: building the "graph" executes it eagerly on the fed inputs -- every transpose / reshape / concat of
`chebyshev5`, the layer wiring of encoder / decoder / discriminator, the variable names and shapes, `loss()` and the
optimiser wiring of `training()` (including its quirks) are the reference's own code.  What stays unpinned is only what
this file states about the primitives themselves (a matmul is a matmul; `leaky_relu` has alpha 0.2; `Reduction.MEAN`
divides by the number of elements; `MomentumOptimizer` is accum = m accum + g, var -= lr accum; ...).

tests/golden/make_ref_golden.py uses it to generate golden vectors from /root/reference; tests/test_reference_golden.py
checks the oracle against them (and, when the reference checkout is present, re-runs the reference).

Graph vs eager: `tf.placeholder(dtype, shape, name)` returns `FEEDS[name]`; variables are created from `PARAMS[name]`
(the reference's TF variable names); `apply_gradients` is DEFERRED until `run_pending()` so that, as in a `sess.run`, every
gradient is taken at the pre-update values.  `tf.random_normal` returns `FEEDS["eps"]` (the VAE noise).
"""
import math
import sys
import types

import numpy as np
import torch

# ---- state of one "graph" -------------------------------------------------------------------------------------------
FEEDS = {}          # placeholder name -> array / tensor
PARAMS = {}         # variable name -> array (initial values)
VARS = {}           # variable name -> TFVar (creation order)
TRAINABLE = []      # names of trainable variables, creation order
REG_LOSSES = []     # (variable name, loss tensor)
RECORD = {}         # what the optimisers saw: "grads" {name: tensor}, "lr" [...], "slots" {name: tensor}
_scope = []         # variable-scope stack: (name, reuse)
_default_names = [{}]
_pending = []
GLOBAL_STEP = None


def reset(feeds=None, params=None, global_step=0, slots=None):
    global GLOBAL_STEP
    FEEDS.clear(); PARAMS.clear(); VARS.clear(); REG_LOSSES.clear(); RECORD.clear()
    del TRAINABLE[:], _scope[:], _pending[:]
    _default_names[:] = [{}]
    FEEDS.update(feeds or {})
    PARAMS.update(params or {})
    GLOBAL_STEP = _GlobalStep(int(global_step))
    RECORD.update(grads={}, lr=[], created=[],
                  slots={k: torch.as_tensor(np.asarray(v, np.float32)).clone() for k, v in (slots or {}).items()})


class _Shape(list):
    def as_list(self):
        return list(self)


class TFVar(torch.Tensor):
    """A trainable leaf with the attributes the reference reads from tf.Variable: `.name` ('scope/weights:0'), `.op.name`."""

    @property
    def name(self):
        return self._tfname + ":0"

    @property
    def op(self):
        return types.SimpleNamespace(name=self._tfname)


class _GlobalStep(object):
    """tf.Variable(0, name='global_step', trainable=False): an integer the learning-rate policy reads."""

    def __init__(self, v):
        self.v = int(v)

    def __sub__(self, o):
        return _GlobalStep(self.v - int(o))

    def __lt__(self, o):
        return self.v < int(o)

    def __int__(self):
        return self.v


def _t(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, _GlobalStep):
        return torch.tensor(x.v, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ---- the module -----------------------------------------------------------------------------------------------------
tf = types.ModuleType("tensorflow")
tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = torch.float32, torch.float64, torch.int32, torch.int64, torch.bool


class _Ctx(object):
    def __init__(self, enter=None, leave=None):
        self._enter, self._leave = enter, leave

    def __enter__(self):
        if self._enter:
            self._enter()
        return self

    def __exit__(self, *exc):
        if self._leave:
            self._leave()
        return False


class _Graph(object):
    def as_default(self):
        return _Ctx()


tf.Graph = _Graph
tf.name_scope = lambda *a, **k: _Ctx()
tf.control_dependencies = lambda deps: _Ctx()
tf.set_random_seed = lambda seed: None
tf.random = types.SimpleNamespace(set_random_seed=lambda seed: None)
tf.global_variables_initializer = lambda: None


def _variable_scope(name, reuse=None, **kw):
    def enter():
        inherited = _scope[-1][1] if _scope else False
        _scope.append((name, bool(reuse) or inherited))
        _default_names.append({})

    def leave():
        _scope.pop()
        _default_names.pop()

    return _Ctx(enter, leave)


tf.variable_scope = _variable_scope


def _full(name):
    return "/".join([s for s, _ in _scope] + [name])


def _get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    full = _full(name)
    if full in VARS:
        if not (_scope and _scope[-1][1]):
            raise ValueError("Variable %s already exists (reuse not set)" % full)       # TF's own error
        return VARS[full]
    if _scope and _scope[-1][1]:
        raise ValueError("Variable %s does not exist (reuse is set)" % full)
    if full in PARAMS:
        val = np.asarray(PARAMS[full], np.float32)
        if shape is not None and tuple(int(s) for s in shape) != tuple(val.shape):
            raise ValueError("variable %s: the reference wants shape %s, the parameter set has %s"
                             % (full, tuple(shape), val.shape))
    else:
        raise KeyError("the reference creates variable %r, which the parameter set does not have" % full)
    v = torch.Tensor._make_subclass(TFVar, torch.from_numpy(val.copy()), bool(trainable))
    v._tfname = full
    VARS[full] = v
    RECORD["created"].append((full, tuple(val.shape)))
    if trainable:
        TRAINABLE.append(full)
    return v


tf.get_variable = _get_variable
tf.trainable_variables = lambda: [VARS[n] for n in TRAINABLE]
tf.constant_initializer = lambda v=0.0, **k: ("constant", v)
tf.truncated_normal_initializer = lambda mean=0.0, stddev=1.0, **k: ("truncated_normal", mean, stddev)
tf.Variable = lambda initial_value=0, name=None, trainable=True, **k: GLOBAL_STEP     # only global_step is built this way


def _placeholder(dtype, shape=None, name=None):
    if name == "is_training":
        return False
    if name not in FEEDS:
        raise KeyError("placeholder %r is not fed" % name)
    x = _t(FEEDS[name], dtype)
    if shape is not None and tuple(int(s) for s in shape) != tuple(x.shape):
        raise ValueError("placeholder %s: shape %s fed, %s declared" % (name, tuple(x.shape), tuple(shape)))
    return x


tf.placeholder = _placeholder
tf.random_normal = lambda shape, mean=0.0, stddev=1.0, dtype=None, **k: _t(FEEDS["eps"]).reshape([int(s) for s in shape])

# ---- dense algebra ---------------------------------------------------------------------------------------------------
tf.reshape = lambda x, shape, name=None: torch.reshape(x, [int(s) for s in shape])
tf.transpose = lambda x, perm=None, name=None: x.permute(*perm) if perm is not None else x.t()
tf.concat = lambda values, axis, name=None: torch.cat([_t(v) for v in values], dim=axis)
tf.stack = lambda values, axis=0, name=None: torch.stack(list(values), dim=axis)
tf.expand_dims = lambda x, axis, name=None: torch.unsqueeze(x, axis)
tf.matmul = lambda a, b, name=None: torch.matmul(a, b)
tf.ones = lambda shape, dtype=None, name=None: torch.ones([int(s) for s in shape])
tf.zeros = lambda shape, dtype=None, name=None: torch.zeros([int(s) for s in shape])
tf.ones_like = lambda x, **k: torch.ones_like(x)
tf.zeros_like = lambda x, **k: torch.zeros_like(x)
tf.identity = lambda x, name=None: x
tf.add = lambda a, b, name=None: a + b
tf.multiply = lambda a, b, name=None: a * b
tf.divide = lambda a, b, name=None: a / b
tf.square = lambda x, name=None: x * x
tf.exp = lambda x, name=None: torch.exp(x)
tf.sqrt = lambda x, name=None: torch.sqrt(x)
tf.abs = lambda x, name=None: torch.abs(x)
tf.equal = lambda a, b, name=None: a == b
tf.shape = lambda x, name=None: list(x.shape)
tf.cast = lambda x, dtype, name=None: _t(x, dtype).to(dtype)
tf.cond = lambda pred, true_fn, false_fn, **k: true_fn() if bool(pred) else false_fn()
tf.where = lambda c, *a, **k: torch.nonzero(c) if not a else torch.where(c, *a)


def _reduce(fn):
    def f(x, axis=None, keepdims=False, keep_dims=False, name=None):
        kd = keepdims or keep_dims
        return fn(x) if axis is None else fn(x, dim=axis, keepdim=kd)
    return f


tf.reduce_mean = _reduce(torch.mean)
tf.reduce_sum = _reduce(torch.sum)


def _gather(params, indices, axis=0, name=None):
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    return torch.index_select(params, axis, idx)


tf.gather = _gather


def _norm(x, ord="euclidean", axis=None, keepdims=False, name=None):
    assert ord in ("euclidean", 2)
    return torch.sqrt(torch.sum(x * x, dim=axis, keepdim=keepdims)) if axis is not None else torch.sqrt(torch.sum(x * x))


tf.norm = _norm


# ---- sparse ------------------------------------------------------------------------------------------------------------
class _Sparse(object):
    def __init__(self, indices, values, dense_shape):
        idx = torch.as_tensor(np.asarray(indices).T.copy(), dtype=torch.long)
        val = torch.as_tensor(np.asarray(values), dtype=torch.float32)          # TF: SparseTensor of the matrix' dtype (fp32)
        self.m = torch.sparse_coo_tensor(idx, val, tuple(int(s) for s in dense_shape)).coalesce()


tf.SparseTensor = _Sparse
tf.sparse_reorder = lambda sp: sp                                               # coalesce() already orders the indices
tf.sparse_tensor_dense_matmul = lambda sp, x, **k: torch.sparse.mm(sp.m, x)

# ---- nn / layers / losses ---------------------------------------------------------------------------------------------
def _moments(x, axes, keep_dims):
    """tf.nn.moments as TF-1.13 computes it: mean, then variance = mean(squared_difference(x, stop_gradient(mean)))."""
    mean = torch.mean(x, dim=axes, keepdim=True)
    var = torch.mean((x - mean.detach()) ** 2, dim=axes, keepdim=True)
    if not keep_dims:
        mean, var = mean.squeeze(axes), var.squeeze(axes)
    return mean, var


tf.nn = types.SimpleNamespace(
    relu=lambda x, name=None: torch.relu(x),
    tanh=lambda x, name=None: torch.tanh(x),
    leaky_relu=lambda x, alpha=0.2, name=None: torch.where(x > 0, x, alpha * x),
    # max(x, 0) - x z + log(1 + exp(-|x|))  (the formula in the op's documentation)
    sigmoid_cross_entropy_with_logits=lambda logits=None, labels=None, name=None: (
        torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-torch.abs(logits)))),
    moments=lambda x, axes, keep_dims=False, **k: _moments(x, axes, keep_dims),
)


def _dense(inputs, units, activation=None, kernel_regularizer=None, trainable=True, name=None, **kw):
    """tf.layers.dense: variables `<scope>/<dense[_k]>/kernel` (glorot-uniform) and `/bias` (zeros); a regularisation
    loss is registered when the kernel is CREATED (not when it is reused)."""
    counts = _default_names[-1]
    base = name or "dense"
    k = counts.get(base, 0)
    counts[base] = k + 1
    layer = base if (name or k == 0) else "%s_%d" % (base, k)
    with _variable_scope(layer):
        existed = _full("kernel") in VARS
        W = _get_variable("kernel", [int(inputs.shape[-1]), int(units)], trainable=trainable)
        b = _get_variable("bias", [int(units)], trainable=trainable)
    if kernel_regularizer is not None and not existed:
        REG_LOSSES.append((W._tfname, kernel_regularizer(W)))
    y = torch.matmul(inputs, W) + b
    return activation(y) if activation is not None else y


tf.layers = types.SimpleNamespace(dense=_dense)
tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
    l2_regularizer=lambda scale, scope=None: (lambda w: scale * torch.sum(w * w) / 2.0),     # scale * tf.nn.l2_loss(w)
    batch_norm=None))


def _get_regularization_loss(scope=None, name=None):
    sel = [l for n, l in REG_LOSSES if scope is None or n.startswith(scope)]
    return sum(sel) if sel else torch.zeros(())


def _weighted_mean(values, weights):
    """tf.losses.compute_weighted_loss with Reduction.MEAN: sum(values * weights) / sum(broadcast weights)."""
    w = _t(weights)
    return torch.sum(values * w) / torch.sum(torch.ones_like(values) * w)


def _huber(labels, predictions, weights=1.0, delta=1.0, reduction=None, **k):
    e = torch.abs(predictions - labels)
    q = torch.clamp(e, max=delta)
    return _weighted_mean(0.5 * q * q + delta * (e - q), weights)


tf.losses = types.SimpleNamespace(
    Reduction=types.SimpleNamespace(MEAN="weighted_mean"),
    get_regularization_loss=_get_regularization_loss,
    absolute_difference=lambda labels=None, predictions=None, weights=1.0, reduction=None, **k:
        _weighted_mean(torch.abs(predictions - labels), weights),
    mean_squared_error=lambda labels=None, predictions=None, weights=1.0, reduction=None, **k:
        _weighted_mean((predictions - labels) ** 2, weights),
    huber_loss=_huber)

# ---- summaries, sessions ------------------------------------------------------------------------------------------------
class _Null(object):
    """Accepts any attribute access / call and does nothing (summary writers, protobuf summaries, savers)."""

    def __getattr__(self, name):
        return _Null()

    def __call__(self, *a, **k):
        return _Null()


# The graph has already been executed when build_graph returns, so there is nothing for a session to run.  Tests that
# want the reference's host loops (fit / predict: what they feed, how often they run) set SESSION_FACTORY to a class
# whose instances record `run(fetches, feed_dict)` calls and answer with stand-in values.
SESSION_FACTORY = None
tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None,
                                   merge_all=lambda *a, **k: "op_summary", FileWriter=lambda *a, **k: _Null())
tf.Summary = lambda *a, **k: _Null()
tf.Session = lambda *a, **k: SESSION_FACTORY(*a, **k) if SESSION_FACTORY is not None else None


class _EMA(object):
    def __init__(self, decay):
        pass

    def apply(self, var_list):
        return None

    def average(self, x):
        return x


# ---- training ------------------------------------------------------------------------------------------------------------
def _exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = int(global_step) / float(decay_steps)
    if staircase:
        p = math.floor(p)
    return _t(learning_rate) * (decay_rate ** p)


def _clip_by_global_norm(t_list, clip_norm, use_norm=None, name=None):
    """scale = clip_norm * min(1 / global_norm, 1 / clip_norm); every tensor times scale."""
    ts = [t for t in t_list if t is not None]
    gn = torch.sqrt(sum(torch.sum(t.detach().double() ** 2) for t in ts)).float()
    scale = clip_norm * torch.minimum(1.0 / gn, torch.tensor(1.0 / clip_norm))
    return [None if t is None else t.detach() * scale for t in t_list], gn


class _Optimizer(object):
    def compute_gradients(self, loss, var_list=None):
        grads = torch.autograd.grad(loss, list(var_list), retain_graph=True, allow_unused=True)
        for g, v in zip(grads, var_list):
            RECORD["grads"][v._tfname] = None if g is None else g.detach().clone().as_subclass(torch.Tensor)
        return list(zip(grads, var_list))

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = [(g, v) for g, v in grads_and_vars]
        _pending.append((self, gv, global_step))
        return None


class _Momentum(_Optimizer):
    """tf.train.MomentumOptimizer (use_nesterov=False): accum = momentum * accum + grad; var -= lr * accum."""

    def __init__(self, learning_rate, momentum, **k):
        self.lr, self.momentum = learning_rate, momentum

    def _apply(self, g, v):
        acc = RECORD["slots"].setdefault(v._tfname + "/Momentum", torch.zeros_like(v.detach()).as_subclass(torch.Tensor))
        acc.mul_(self.momentum).add_(g.as_subclass(torch.Tensor))
        v.data.sub_(_t(self.lr).float() * acc)


class _Adam(_Optimizer):
    """tf.train.AdamOptimizer (beta1 0.9, beta2 0.999, epsilon 1e-8 as float32): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);
    m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; var -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **k):
        self.lr, self.b1, self.b2, self.eps, self.t = learning_rate, float(np.float32(beta1)), float(np.float32(beta2)), epsilon, 0

    def _apply(self, g, v):
        z = lambda: torch.zeros_like(v.detach()).as_subclass(torch.Tensor)
        m = RECORD["slots"].setdefault(v._tfname + "/Adam", z())
        s = RECORD["slots"].setdefault(v._tfname + "/Adam_1", z())
        g = g.as_subclass(torch.Tensor)
        m.mul_(self.b1).add_((1 - self.b1) * g)
        s.mul_(self.b2).add_((1 - self.b2) * g * g)
        lr_t = _t(self.lr).float() * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        v.data.sub_(lr_t * m / (torch.sqrt(s) + self.eps))


def run_pending():
    """What `sess.run([op_train_g, op_train_d])` does after the gradients: both apply ops, each bumping global_step."""
    for opt, gv, step in _pending:
        RECORD["lr"].append(float(_t(opt.lr)))
        if isinstance(opt, _Adam):
            opt.t += 1
        with torch.no_grad():
            for g, v in gv:
                if g is not None:
                    opt._apply(g.detach(), v)
        if step is not None:
            step.v += 1
    del _pending[:]


tf.train = types.SimpleNamespace(
    exponential_decay=_exponential_decay, MomentumOptimizer=_Momentum, AdamOptimizer=_Adam,
    ExponentialMovingAverage=_EMA, Saver=lambda *a, **k: _Null(), latest_checkpoint=lambda *a, **k: None)
tf.clip_by_global_norm = _clip_by_global_norm


# ---- installation ---------------------------------------------------------------------------------------------------------
def install(template_vertices=None):
    """Put the shim (and stubs of the reference's other unavailable imports) into sys.modules; give torch tensors the two
    tf.Tensor methods the reference calls (`get_shape()`, with `.as_list()`)."""
    torch.Tensor.get_shape = lambda self: _Shape(int(s) for s in self.shape)
    sys.modules["tensorflow"] = tf
    py = types.ModuleType("tensorflow.python")
    util = types.ModuleType("tensorflow.python.util")
    util.deprecation = types.SimpleNamespace(_PRINT_DEPRECATION_WARNINGS=False)
    py.util = util
    tf.python = py
    sys.modules["tensorflow.python"] = py
    sys.modules["tensorflow.python.util"] = util
    # lib/models.py: `trimesh.load(template).vertices` is added to predictions and targets in the edge loss (it cancels);
    # TF would convert the float64 array to a float32 tensor there, so the stub hands out a float32 tensor
    tm = types.ModuleType("trimesh")
    verts = None if template_vertices is None else torch.as_tensor(np.asarray(template_vertices), dtype=torch.float32)
    tm.load = lambda *a, **k: types.SimpleNamespace(vertices=verts)
    sys.modules.setdefault("trimesh", tm)
    sys.modules.setdefault("smplx", types.ModuleType("smplx"))                 # demos.py imports it for demo_full only
    if "psbody" not in sys.modules:                                             # lib/load_data.py: Mesh(filename=...) only
        ps, pm = types.ModuleType("psbody"), types.ModuleType("psbody.mesh")
        pm.Mesh = lambda filename=None, v=None, f=None: types.SimpleNamespace(filename=filename, v=v, f=f)
        ps.mesh = pm
        sys.modules["psbody"], sys.modules["psbody.mesh"] = ps, pm
    for name in ("cv2", "matplotlib", "matplotlib.pyplot"):                     # imported by lib/utils.py, unused on this path
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules.get("matplotlib.pyplot")
    return tf
