import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_training_host import load_train_case
from turtlevsr_b200.training import TrainStep
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
OPTIM = dict(type="Adam", lr=4e-4, weight_decay=0, betas=[0.9, 0.99])
for case in ["train_tiny_t0.npz", "train_tiny_t1.npz"]:
    net, lq, gt, z = load_train_case(case)
    net = net.cuda(); ts = TrainStep(net, OPTIM); lq, gt = lq.cuda(), gt.cuda()
    l1 = ts.step(lq, gt).item(); l2 = ts.step(lq, gt).item()
    worst, n_off, n_all = 0.0, 0, 0
    for name, p in net.named_parameters():
        want = torch.from_numpy(z["after2::" + name]).cuda()
        d = (p.detach() - want).abs()
        worst = max(worst, d.max().item()); n_off += int((d > 2e-5).sum()); n_all += d.numel()
    print(case, "dl1", abs(l1 - float(z["losses"][0])), "dl2", abs(l2 - float(z["losses"][1])), "worst", worst, "frac_off", n_off / n_all)
