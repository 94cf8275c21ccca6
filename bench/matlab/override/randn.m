function r = randn(varargin)
%RANDN Shadow of the built-in, active only while bench/matlab/dump_reference_vectors.m has this folder on the path.
%   ent/entmc_vbmc.m:53 draws  epsilon = randn(D,1,Ns/2)  once per mixture component j = 1..K, in that order.  The
%   committed draws are stored as EPS(D,Ns/2,K) (component j = page j); each call returns the next page, reshaped to
%   the requested size.  Any other call pattern is an error: nothing else on the dumped path may consume random numbers.
global VBMC_B200_EPS VBMC_B200_EPS_NEXT
if isempty(VBMC_B200_EPS)
    error('vbmc_b200:randn','randn override active but no draws were queued.');
end
sz = cell2mat(varargin);
j = VBMC_B200_EPS_NEXT;
if j > size(VBMC_B200_EPS,3)
    error('vbmc_b200:randn','more randn calls than queued components (%d).', size(VBMC_B200_EPS,3));
end
page = VBMC_B200_EPS(:,:,j);                 % D x Ns/2
if numel(sz) ~= 3 || sz(1) ~= size(page,1) || sz(2) ~= 1 || sz(3) ~= size(page,2)
    error('vbmc_b200:randn','unexpected randn size [%s]; entmc_vbmc asks for (D,1,Ns/2).', num2str(sz));
end
r = reshape(page,[sz(1),1,sz(3)]);
VBMC_B200_EPS_NEXT = j + 1;
end
