import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import standin_model, load_golden, rel_l2
T = 20
for name in ['nu_like', 'msr80c', 'co', 'msr3c']:
    g = load_golden(f'standin_{name}.npz')
    for prec in ['fp16x3', 'fp16x2']:
        ddpm, cfg = standin_model(name, 'cuda:0')
        ddpm.model.precision = prec
        c = lambda a: torch.as_tensor(a).to('cuda:0')
        with torch.no_grad():
            eps = ddpm.model(c(g['x']), c(g['ts']) / T, c(g['cond']), c(g['mask']))
        torch.cuda.synchronize()
        print(f'{name:8s} {prec}: forward rel-L2 {rel_l2(eps.cpu(), g["eps"]):.3e}', flush=True)
        y0 = ddpm.sample(c(g['cond']), 3.0, y_init=c(g['y_T']), noise=c(g['noise']))
        torch.cuda.synchronize()
        print(f'{name:8s} {prec}: sampler(omega=3) rel-L2 {rel_l2(y0.cpu(), g["y0_omega3"]):.3e}', flush=True)
