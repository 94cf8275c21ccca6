function dump_reference_vectors(cases)
%DUMP_REFERENCE_VECTORS Run the unmodified VBMC reference on the committed inputs and store its outputs.
%   Reads  tests/golden/matlab_inputs/<case>.mat  (tests/golden/export_inputs_mat.py) and writes
%   tests/golden/reference_<case>.mat, which tests/test_reference_vectors.py compares with the oracle and the CUDA path.
%   The reference must be on the path (addpath(genpath(<vbmc>))); private functions are reached through their public callers.
here = fileparts(mfilename('fullpath'));
root = fullfile(here,'..','..');
indir = fullfile(root,'tests','golden','matlab_inputs');
if nargin < 1 || isempty(cases)
    d = dir(fullfile(indir,'*.mat'));
    cases = arrayfun(@(f) f.name(1:end-4), d, 'UniformOutput', false);
end
if ischar(cases); cases = {cases}; end
global VBMC_B200_EPS VBMC_B200_EPS_NEXT
for ic = 1:numel(cases)
    in = load(fullfile(indir,[cases{ic} '.mat']));
    D = double(in.D); K = double(in.K); Ns = double(in.Ns);

    % ---- GP posterior with the reference's own gplite_post (gplite/gplite_post.m) ----
    s2 = []; noisefun = [1 0 0];
    if isfield(in,'s2') && ~isempty(in.s2); s2 = in.s2(:); noisefun = [1 1 0]; end
    gp = gplite_post(in.hyp, in.X, in.y(:), 1, double(in.meanfun), noisefun, s2);

    % ---- variational posterior (fields of misc/setupvars_vbmc.m:78-99 that the path reads) ----
    vp.D = D; vp.K = K;
    vp.mu = in.mu;                       % D x K
    vp.sigma = in.sigma(:)';             % 1 x K
    vp.lambda = in.lambda(:);            % D x 1
    vp.w = in.w(:)';                     % 1 x K
    vp.eta = in.eta(:)';                 % 1 x K
    vp.optimize_mu = true; vp.optimize_sigma = true; vp.optimize_lambda = true; vp.optimize_weights = true;
    vp.delta = [];
    options.TolLength = in.TolLength; options.TolWeight = in.TolWeight;
    options.TolConLoss = in.TolConLoss; options.WeightPenalty = in.WeightPenalty;
    [vp,thetabnd] = vpbounds(vp,gp,options,K);
    theta = in.theta(:);

    out = struct();
    out.alpha = zeros(size(in.X,1), numel(gp.post));
    for s = 1:numel(gp.post); out.alpha(:,s) = gp.post(s).alpha; end
    out.thetabnd_lb = thetabnd.lb; out.thetabnd_ub = thetabnd.ub;

    % ---- negelcbo_vbmc with the committed draws in place of the global randn stream ----
    addpath(fullfile(here,'override'));
    cleanup = onCleanup(@() rmpath(fullfile(here,'override')));
    VBMC_B200_EPS = in.epsilon; VBMC_B200_EPS_NEXT = 1;                      % D x Ns/2 x K
    [out.F,out.dF,out.G,out.H,~,out.dH] = negelcbo_vbmc(theta,0,vp,gp,Ns,1,0,0,thetabnd,0);
    % the same call unpacked: vp as negelcbo_vbmc sees it after reading theta (negelcbo_vbmc.m:32-48)
    vpt = vp;
    vpt.mu(:,:) = reshape(theta(1:D*K),[D,K]);
    vpt.sigma(1,:) = exp(theta(D*K+(1:K)));
    vpt.lambda(:,1) = exp(theta(D*K+K+(1:D)));
    vpt.eta(1,:) = theta(end-K+1:end); vpt.w(1,:) = exp(vpt.eta); vpt.w = vpt.w/sum(vpt.w);
    VBMC_B200_EPS_NEXT = 1;
    [out.H_entmc,out.dH_entmc] = entmc_vbmc(vpt,Ns,[1 1 1 1],1);
    clear cleanup                                                            % built-in randn again

    % ---- deterministic pieces ----
    [out.G_glj,out.dG_glj,out.varG_diag,out.dvarG_diag,out.varss_diag] = gplogjoint(vpt,gp,[1 1 1 1],1,1,2);   % diagonal variance + its gradient
    [~,~,out.varG_full,~,out.varss_full,out.I_sk,out.J_sjk] = gplogjoint(vpt,gp,0,1,1,1,1);                     % full variance, per-component terms
    [out.H_lb,out.dH_lb] = entlb_vbmc(vpt,[1 1 1 1],1);
    [out.F_lb,out.dF_lb] = negelcbo_vbmc(theta,0,vp,gp,0,1,0,0,thetabnd,0);
    gp1 = gplite_post(in.hyp(:,1), in.X, in.y(:), 1, double(in.meanfun), noisefun, s2);
    [out.nlZ,out.dnlZ] = gplite_nlZ(in.hyp(:,1),gp1,[]);
    [out.ymu,out.ys2,out.fmu,out.fs2] = gplite_pred(gp,in.Xstar);

    save(fullfile(root,'tests','golden',['reference_' cases{ic} '.mat']),'-struct','out','-v7');
    fprintf('%s: F = %.16g, H = %.16g, G = %.16g\n', cases{ic}, out.F, out.H, out.G);
end
VBMC_B200_EPS = []; VBMC_B200_EPS_NEXT = [];
end
