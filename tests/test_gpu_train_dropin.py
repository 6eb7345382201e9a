"""End-to-end drop-in: the reference's OWN, unmodified train.py (train.py:20-97) + checkpoint_utils.py:20-81 + utils.py +
data/dataloaders.py driving plainlm_b200 through the dropin/ shims, next to the same script driving the unmodified
reference (eager PyTorch on the same GPU) on the same on-disk dataset, config and seed.

  ours      : PYTHONSAFEPATH=1 PYTHONPATH=dropin:<repo>:baseline/_ref  python baseline/_ref/train.py --config=...
  reference : PYTHONSAFEPATH=1 PYTHONPATH=baseline/_ref                python baseline/_ref/train.py --config=...

(PYTHONSAFEPATH keeps the script's own directory — the reference root, with its models/ engine/ optim/ — off the front
of sys.path, so PYTHONPATH order decides which implementation train.py imports.)  Checked: the per-step loss curves
agree within 1 % (BASELINE.json north_star), both write a checkpoint through the reference's save_checkpoint, and each
implementation resumes from the OTHER's checkpoint (cross-loading) and continues on the same curve.

baseline/_ref is the git-ignored install of the unmodified reference (baseline/install_ref.sh); it travels to the GPU box
with the gpurun snapshot.  Nothing here reads /root/reference."""

import os
import re
import subprocess
import sys

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
T, V, B, ACCUM = 128, 512, 4, 2


def _markov_rows(n_rows, seed=11):
  """Low-entropy synthetic stream: order-1 Markov chain over a 64-symbol subset of the vocabulary, so
  the loss falls quickly from ln V and a 1 % curve comparison means something."""
  g = torch.Generator().manual_seed(seed)
  sym = torch.randperm(V, generator=g)[:64]
  trans = torch.softmax(torch.randn(64, 64, generator=g) * 2.5, dim=1)
  rows = torch.empty(n_rows, T + 1, dtype=torch.int64)
  state = torch.randint(0, 64, (n_rows,), generator=g)
  for t in range(T + 1):
    rows[:, t] = sym[state]
    state = torch.multinomial(trans[state], 1, generator=g).squeeze(1)
  return rows


def _config(tmp, data_dir, **over):
  cfg = dict(
    deterministic=False, seed=100, trainset_path=data_dir, vocab_size=V, seq_len=T, sampler='sequential', sampler_seed=None,
    num_workers=0, eval=False, validset_path=None, eval_every_steps=None, model='transformer', d_model=256, mlp_class='glu',
    expand='8/3', n_layers=2, n_heads=4, rms_norm=True, tie_embeddings=False, torch_compile=False, steps_budget=12,
    micro_batch_size=B, grad_accumulation_steps=ACCUM, dtype='bfloat16', optim='adamw', fused_optim=True, lr=3e-3,
    weight_decay=0.1, beta1=0.9, beta2=0.95, grad_clip=1.0, scheduler='warmup_cosine', warmup_steps=0.25,
    cooldown_steps=None, lr_start=0.0, lr_end=1e-5, lr_end_pct=None, log_every_steps=1, print_progress=True,
    use_wandb=False, wandb_project='x', wandb_dir=str(tmp), wandb_run_name='x', exp_name='run', out_dir=str(tmp),
    over_write=True, resume=False, resume_step=None, resume_exp_name=None, save_last_checkpoint=True,
    save_intermediate_checkpoints=False, save_every_steps=None, intra_doc_masking=False)
  cfg.update(over)
  path = os.path.join(str(tmp), f"{cfg['exp_name']}{'_resume' if cfg['resume'] else ''}.yaml")
  with open(path, 'w') as f:
    yaml.safe_dump(cfg, f)
  return path


def _run_train(cfg_path, ours):
  env = dict(os.environ, PYTHONSAFEPATH='1', TORCHDYNAMO_DISABLE='1', WANDB_MODE='disabled', CUDA_VISIBLE_DEVICES='0')
  for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
    env.pop(k, None)
  env['PYTHONPATH'] = os.pathsep.join(([os.path.join(ROOT, 'dropin'), ROOT] if ours else []) + [REF])
  p = subprocess.run([sys.executable, os.path.join(REF, 'train.py'), f'--config={cfg_path}'], env=env, cwd=ROOT,
                     capture_output=True, text=True, timeout=900)
  assert p.returncode == 0, f'train.py failed (ours={ours}):\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}'
  assert '=== Training Completed! ===' in p.stdout
  steps, losses = [], []
  for line in p.stdout.splitlines():
    m = re.search(r'\bstep: (\d+) \|.*train/loss: ([0-9.e+-]+)', line)
    if m:
      steps.append(int(m.group(1)))
      losses.append(float(m.group(2)))
  return steps, losses, p.stdout


@pytest.fixture(scope='module')
def workdir(tmp_path_factory):
  if not os.path.isfile(os.path.join(REF, 'train.py')):
    pytest.skip('baseline/_ref/train.py missing: run baseline/install_ref.sh where /root/reference exists (build() does)')
  import datasets

  tmp = tmp_path_factory.mktemp('dropin')
  rows = _markov_rows(24 * ACCUM * B)
  data_dir = os.path.join(str(tmp), 'train')
  datasets.Dataset.from_dict({'input_ids': rows.tolist()}).with_format('torch').save_to_disk(data_dir)
  return tmp, data_dir


def test_reference_train_py_runs_on_the_dropin_and_matches_the_reference(workdir):
  tmp, data_dir = workdir
  s_ours, l_ours, _ = _run_train(_config(tmp, data_dir, exp_name='ours'), ours=True)
  s_ref, l_ref, _ = _run_train(_config(tmp, data_dir, exp_name='ref'), ours=False)
  assert s_ours == s_ref and len(l_ours) >= 12, (s_ours, s_ref)
  assert l_ours[0] > 5.5 and l_ours[-1] < l_ours[0] - 1.0, f'loss did not fall: {l_ours}'
  for s, a, b in zip(s_ours, l_ours, l_ref):
    assert abs(a - b) <= 1e-2 * abs(b), f'step {s}: ours {a} vs reference {b}\nours {l_ours}\nref  {l_ref}'
  # both wrote a checkpoint through the reference's checkpoint_utils.save_checkpoint with the reference's keys
  # (train.py:69,92-94 leaves the loop at step budget + 1, so the file is ckpt_step_13 after 12 optimizer steps)
  for name in ('ours', 'ref'):
    ck = torch.load(os.path.join(str(tmp), name, 'ckpt_step_13.pth'), map_location='cpu')
    assert set(ck) == {'step', 'state_dict', 'optimizer', 'scheduler', 'scaler'} and ck['step'] == 13
  ko = torch.load(os.path.join(str(tmp), 'ours', 'ckpt_step_13.pth'), map_location='cpu')
  kr = torch.load(os.path.join(str(tmp), 'ref', 'ckpt_step_13.pth'), map_location='cpu')
  assert list(ko['state_dict']) == list(kr['state_dict'])
  assert {k: v.shape for k, v in ko['state_dict'].items()} == {k: v.shape for k, v in kr['state_dict'].items()}
  st_o = ko['optimizer']['state'][0]
  st_r = kr['optimizer']['state'][0]
  assert set(st_o) == set(st_r) == {'step', 'exp_avg', 'exp_avg_sq'}
  # weights after 12 optimizer steps: same trajectory within bf16 training noise
  for k in ko['state_dict']:
    a, b = ko['state_dict'][k].float(), kr['state_dict'][k].float()
    assert (a - b).norm() <= 0.05 * b.norm() + 1e-6, k


def test_each_implementation_resumes_from_the_others_checkpoint(workdir):
  tmp, data_dir = workdir
  for name in ('ours', 'ref'):
    if not os.path.exists(os.path.join(str(tmp), name, 'ckpt_step_13.pth')):
      pytest.skip('needs the checkpoints written by the previous test')
  common = dict(resume=True, resume_step=13, steps_budget=17, save_last_checkpoint=False)
  # ours <- reference-written checkpoint; reference <- ours-written checkpoint (maybe_load_checkpoint, engine resume path)
  s_a, l_a, _ = _run_train(_config(tmp, data_dir, exp_name='ours_from_ref', resume_exp_name='ref', **common), ours=True)
  s_b, l_b, _ = _run_train(_config(tmp, data_dir, exp_name='ref_from_ours', resume_exp_name='ours', **common), ours=False)
  assert s_a == s_b and len(l_a) >= 4 and s_a[0] == 14, (s_a, s_b)
  for s, a, b in zip(s_a, l_a, l_b):
    assert abs(a - b) <= 1.5e-2 * abs(b), f'resumed step {s}: ours<-ref {a} vs ref<-ours {b}'
