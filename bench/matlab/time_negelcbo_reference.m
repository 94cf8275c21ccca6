function time_negelcbo_reference(casename, nsteps)
%TIME_NEGELCBO_REFERENCE negelcbo_vbmc gradient-steps/s of the unmodified reference (BASELINE.json's CPU figure).
%   Inputs as for DUMP_REFERENCE_VECTORS (use  export_inputs_mat.py --full  for c3: D=10, N=2000, K=50, Ns=32768, S=20).
%   The draws come from MATLAB's own randn here — this measures the reference as it runs in production.
if nargin < 2; nsteps = 20; end
here = fileparts(mfilename('fullpath'));
in = load(fullfile(here,'..','..','tests','golden','matlab_inputs',[casename '.mat']));
D = double(in.D); K = double(in.K); Ns = double(in.Ns);
s2 = []; noisefun = [1 0 0];
if isfield(in,'s2') && ~isempty(in.s2); s2 = in.s2(:); noisefun = [1 1 0]; end
gp = gplite_post(in.hyp, in.X, in.y(:), 1, double(in.meanfun), noisefun, s2);
vp.D = D; vp.K = K; vp.mu = in.mu; vp.sigma = in.sigma(:)'; vp.lambda = in.lambda(:); vp.w = in.w(:)'; vp.eta = in.eta(:)';
vp.optimize_mu = true; vp.optimize_sigma = true; vp.optimize_lambda = true; vp.optimize_weights = true; vp.delta = [];
options.TolLength = in.TolLength; options.TolWeight = in.TolWeight; options.TolConLoss = in.TolConLoss; options.WeightPenalty = in.WeightPenalty;
[vp,thetabnd] = vpbounds(vp,gp,options,K);
theta = in.theta(:);
[F,dF] = negelcbo_vbmc(theta,0,vp,gp,Ns,1,0,0,thetabnd,0); %#ok<ASGLU> warm-up
m = zeros(size(theta)); v = m; t0 = tic;
for it = 1:nsteps
    [F,dF] = negelcbo_vbmc(theta,0,vp,gp,Ns,1,0,0,thetabnd,0); %#ok<ASGLU>
    m = 0.9*m + 0.1*dF; v = 0.999*v + 0.001*dF.^2;                          % utils/fminadam.m:51-60
    step = 0.001 + (0.1-0.001)*exp(-it/200);
    theta = theta - step*(m/(1-0.9^it))./(sqrt(v/(1-0.999^it)) + sqrt(eps));
end
dt = toc(t0);
fprintf('%s: %d steps in %.3f s = %.3f grad-steps/s on %d computational threads\n', casename, nsteps, dt, nsteps/dt, maxNumCompThreads);
end
